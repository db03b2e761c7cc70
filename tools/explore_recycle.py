"""GPU exploration (scratch): per-point cost of a sweep with subspace recycling."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from explore_scale import make
from emerge_b200.sweep import FrequencySweep, hierarchical_order

nx, ny, nz = [int(v) for v in sys.argv[1:4]]
npts = int(sys.argv[4]); cap = int(sys.argv[5]); mode = sys.argv[6] if len(sys.argv) > 6 else "hier"
limit = int(sys.argv[7]) if len(sys.argv) > 7 else npts
freqs = np.linspace(8e9, 12e9, npts)
t, er, ur, bcs, dims = make(nx, ny, nz)
sw = FrequencySweep(t, er, ur, bcs, recycle=cap)
sw.setup()
for p in sw.ports: p.active = False
order = hierarchical_order(npts) if mode == "hier" else list(range(npts))
T0 = time.perf_counter(); tot_it = 0
for n, i in enumerate(order[:limit]):
    f = freqs[i]
    t0 = time.perf_counter()
    sw.ctx.assemble_KM()
    S, st, _ = sw.solve_point(f, raise_on_fail=False)
    dt = time.perf_counter() - t0
    tot_it += sum(s['iters'] for s in st)
    if n < 40 or n % 20 == 0 or any(s['iters'] for s in st):
        print(f"{n:3d} i={i:3d} f={f/1e9:.3f} {dt*1e3:8.1f} ms  iters {[s['iters'] for s in st]} proj {[('%.1e' % s.get('proj_relres', -1)) for s in st]} "
              f"relres {[('%.1e' % s['relres']) for s in st]} n={st[-1].get('recycled')} |S21|={abs(S[1,0]):.6f} ang={np.angle(S[1,0],deg=True):.3f}", flush=True)
tot = time.perf_counter() - T0
print(f"TOTAL {limit} points in {tot:.1f} s = {limit/tot:.3f} points/s, total iterations {tot_it}, recycle spmvs {sw.ctx.recycle_info()}")
