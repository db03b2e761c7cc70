"""Host wall time per C-ABI entry point inside bench.py's end-to-end leg (first leg of the process, then a second one).
usage: EMB_API_TRACE=1 python tools/e2e_trace.py [K]"""
import os
import sys
import time

os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
os.environ.setdefault("EMB_API_TRACE", "1")
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
from emerge_b200 import lib as L  # noqa: E402
from emerge_b200.sweep import FrequencySweep  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 20


class A:
    recycle, snap, no_coarse_basis, rtol, precond = 40, 0.3, False, 1e-8, "multilevel"


t0 = time.perf_counter()
box, t, er, ur, bcs, Lz = bench.make_waveguide(44, 20, 190, device=0)
print(f"mesh + tables {time.perf_counter()-t0:.2f} s | {L.api_trace_report()}", flush=True)
wbox, wt, wer, wur, wbcs, _ = bench.make_waveguide(8, 4, 12)
w = FrequencySweep(wt, wer, wur, wbcs, device=0, recycle=40)
w.run(list(bench.FREQS[::60]), raise_on_fail=False)
w.ctx.close()
L.api_trace_report()
for rep in range(2):
    r = bench.e2e_pass(A, torch, None, 0, 1, 0, t, er, ur, bcs, K, torch.cuda.synchronize)
    print(f"e2e leg {rep}: {r['ms']:.0f} ms for {r['points']} points; setup {r['setup']}; split {r['split']}", flush=True)
    print("   " + L.api_trace_report(), flush=True)
