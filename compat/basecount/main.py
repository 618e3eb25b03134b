from basecount_b200.main import BaseCount, get_basecounts, handle_arg, run  # noqa: F401
