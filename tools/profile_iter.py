"""Profiling window over the Krylov iterations of one cold solve (run under
`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none`).
usage: profile_iter.py nx,ny,nz [maxit]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from emerge_b200.sweep import FrequencySweep  # noqa: E402

cells = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "24,12,80").split(","))
maxit = int(sys.argv[2]) if len(sys.argv) > 2 else 40
box, t, er, ur, bcs, L = bench.make_waveguide(*cells)
sw = FrequencySweep(t, er, ur, bcs, device=0)
sw.f_ref = 10e9
sw.setup()
sw.solver_opts["maxit"] = maxit
sw.ctx.profiler(True)
S, st, _ = sw.solve_point(10e9, raise_on_fail=False)
sw.ctx.profiler(False)
print("tets", t.tets.shape[1], "stats", st)
