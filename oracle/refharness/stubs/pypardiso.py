"""Stand-in for the (absent) `pypardiso`, used ONLY by the oracle harness (TEST INFRASTRUCTURE).
The reference calls pypardiso.spsolve at fem/solver.py:308; MKL is not installable here, so the
shim forwards to SciPy's SuperLU, the reference's own documented fallback (fem/solver.py:570-573)."""
from scipy.sparse.linalg import spsolve as _sp


def spsolve(A, b, *a, **k):
    return _sp(A.tocsc(), b)
