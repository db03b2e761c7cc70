"""Where does the first end-to-end pass of a process lose time?  Two consecutive e2e passes (fresh sweep object each) of the
first K points with per-point host wall times and the solver's own device times.  usage: e2e_probe.py [K]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
from emerge_b200.sweep import FrequencySweep, hierarchical_order  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 12
box, t, er, ur, bcs, L = bench.make_waveguide(44, 20, 190, device=0)
wbox, wt, wer, wur, wbcs, _ = bench.make_waveguide(8, 4, 12)
w = FrequencySweep(wt, wer, wur, wbcs, device=0, recycle=40)
w.run(list(bench.FREQS[::60]), raise_on_fail=False)
w.ctx.close()
N = t.n_field
outs = {p.port_number: torch.empty(N, dtype=torch.complex128).pin_memory().numpy() for p in bcs[1:]}
order = hierarchical_order(len(bench.FREQS))[:K]
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sw = FrequencySweep(t, er, ur, bcs, device=0, recycle=40)
    sw.f_ref = float(np.median(bench.FREQS))
    sw.setup()
    t1 = time.perf_counter()
    for p in sw.ports:
        p.active = False
    sw.ctx.fields_async(True)
    rows = []
    for i in order:
        a = time.perf_counter()
        S, st, _ = sw.solve_point(bench.FREQS[i], raise_on_fail=False, out_bufs=outs)
        b = time.perf_counter()
        rows.append((i, round((b - a) * 1e3, 1), round(st[0]["ms"], 1), st[0]["iters"]))
    sw.ctx.fields_async(False)
    t2 = time.perf_counter()
    print(f"pass {rep}: setup {t1 - t0:.2f} s {dict(sw.timings)}  points {t2 - t1:.2f} s", flush=True)
    print("   (index, wall ms, solver ms, iters):", rows, flush=True)
    sw.ctx.close()
