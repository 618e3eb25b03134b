"""basecount_b200 -- B200-native pileup counting path, drop-in for tombch/basecount."""
from .version import __version__  # noqa: F401


def __getattr__(name):
    if name == "BaseCount":            # `from basecount import BaseCount` (basecount/__init__.py:1)
        from .main import BaseCount
        return BaseCount
    raise AttributeError(name)
