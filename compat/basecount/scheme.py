from basecount_b200.scheme import load_scheme  # noqa: F401
