"""emerge_b200: B200-native frequency-domain hot path of EMerge."""
import os as _os

# Kernels are loaded when the CUDA context is created, not at their first launch (CUDA's default since 11.7): with lazy
# loading the first sweep of a process lost ~6 s at 1M tets to first launches inside solves and CUDA-graph captures
# (tools/e2e_probe.py).  Only effective when this package is imported before the CUDA driver initialises; an explicit
# CUDA_MODULE_LOADING in the environment wins.
_os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
