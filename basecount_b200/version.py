__version__ = "1.7.2"          # tracks the reference it is a drop-in for (basecount/version.py:1)
