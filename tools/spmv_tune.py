"""Operator-application timing for the current EMB_SPMV_KPR (development probe).  usage: spmv_tune.py nx,ny,nz"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from emerge_b200.sweep import FrequencySweep  # noqa: E402

cells = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "24,12,80").split(","))
box, t, er, ur, bcs, L = bench.make_waveguide(*cells)
sw = FrequencySweep(t, er, ur, bcs, device=0, recycle=0)
sw.solver_opts.update(precond="block")
sw.setup()
ctx = sw.ctx
sw.assemble_frequency(10e9)
nnz, Ns = int(ctx.lib.emb_csr_nnz(ctx.h, 2)), ctx.n_solve
only = os.environ.get("SPMV_TUNE_ONLY")          # e.g. "2,1": nv = 2, complex64 only
cases = [(int(only.split(",")[0]), bool(int(only.split(",")[1])))] if only else [(nv, f) for nv in (1, 2, 4) for f in (False, True)]
for nv, fp32 in cases:
    if True:
        ms = ctx.spmv_bench(20, nv=nv, fp32=fp32)
        b = (8 if fp32 else 16) * nnz + nnz + (4 + 32 * nv) * Ns
        print(f"KPR={os.environ.get('EMB_SPMV_KPR')} SELL={os.environ.get('EMB_SPMV_SELL', '1')} TMA={os.environ.get('EMB_SPMV_TMA', '1')} sigma={os.environ.get('EMB_SELL_SIGMA', 'default')} paired={ctx.paired} nv={nv} fp32={fp32}: "
              f"{ms:.3f} ms  {b / ms / 1e6:.0f} GB/s", flush=True)
