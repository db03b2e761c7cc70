"""Stand-in for the (absent) `numba_progress` package, used ONLY by the oracle harness (TEST INFRASTRUCTURE).

One definition of the no-op progress bar and its numba type lives in `.progress`; the reference imports the class from
here and the type from there (fem/physics/edm/sc.py:18,24), so both names must refer to the same objects.
"""
from .progress import ProgressBar, ProgressBarType, _PBType  # noqa: F401  (_PBType: pickled numba caches name it here)
