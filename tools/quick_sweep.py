"""First K points of the benchmark sweep with per-point statistics (development probe).
usage: quick_sweep.py nx,ny,nz K [lockstep]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from emerge_b200.sweep import FrequencySweep, hierarchical_order  # noqa: E402

cells = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "24,12,80").split(","))
K = int(sys.argv[2]) if len(sys.argv) > 2 else 5
lock = int(sys.argv[3]) if len(sys.argv) > 3 else 4
box, t, er, ur, bcs, L = bench.make_waveguide(*cells)
sw = FrequencySweep(t, er, ur, bcs, device=0)
sw.lockstep = lock
sw.solver_opts.update(rtol=1e-8)
sw.f_ref = float(np.median(bench.FREQS))
t0 = time.perf_counter()
sw.setup()
print("setup s", time.perf_counter() - t0, "tets", t.tets.shape[1], flush=True)
ctx = sw.ctx
sw.assemble_frequency(10e9)
for nv in (1, 2, 4):
    for fp32 in (False, True):
        ms = ctx.spmv_bench(20, nv=nv, fp32=fp32)
        nnz, Ns = int(ctx.lib.emb_csr_nnz(ctx.h, 2)), ctx.n_solve
        b = (9 if fp32 else 17) * nnz + (4 + 32 * nv) * Ns
        print(f"spmv nv={nv} fp32={fp32}: {ms:.3f} ms  {b / ms / 1e6:.0f} GB/s", flush=True)
order = hierarchical_order(len(bench.FREQS))[:K]
tot0 = time.perf_counter()
for i in order:
    ctx.timer_start()
    S, st, _ = sw.solve_point(bench.FREQS[i], raise_on_fail=False)
    ms = ctx.timer_stop()
    sp, spn = ctx.spmv_sampled()
    pp, ppn = ctx.precond_sampled()
    print(json.dumps(dict(i=i, f=bench.FREQS[i], ms=ms, iters=st[0]["iters"], relres=[s["relres"] for s in st],
                          proj=st[0].get("proj_relres"), recycled=st[0].get("recycled"), S21=abs(S[1, 0]), spmv_ms=sp, prec_ms=pp,
                          ms_per_iter=(ms / st[0]["iters"] if st[0]["iters"] else None))), flush=True)
print("total s", time.perf_counter() - tot0, "launches", ctx.launches, "graph iters", ctx.graph_launches)
