"""Plotting stand-in (see oracle/refshim/_null.py): nothing is drawn."""
from _null import NULL as _N


def __getattr__(name):
    if name.startswith("__") and name.endswith("__"):
        raise AttributeError(name)
    return _N
