"""Drop-in for the reference's top-level native module `count` (setup.py:5, count.cpp:102-105):
`from count import bcount` keeps working; the work happens on the GPU."""
from basecount_b200.count import bcount  # noqa: F401
