"""Drop-in package name: `from basecount import BaseCount` (reference basecount/__init__.py:1)."""
from basecount_b200.main import BaseCount  # noqa: F401
from basecount_b200.version import __version__  # noqa: F401
