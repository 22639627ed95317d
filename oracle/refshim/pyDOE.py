"""pyDOE stand-in: `lhs` is the restatement of pyDOE's default algorithm in oracle/data.py, drawing
from NumPy's legacy global RNG like the original (TEST INFRASTRUCTURE ONLY)."""
from oracle.data import lhs  # noqa: F401
