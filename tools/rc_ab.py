"""The reduced-basis bookkeeping on the small recycled sweep of tests/test_gpu_parity.py: accepted points, iterations.
EMB_RC_DEBUG=1 prints, per new direction, |W_t u - Q R_t| / |W_t u| and the orthonormality of the new basis columns."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from emerge_b200.sweep import FrequencySweep
from emerge_b200.synthmesh import box_mesh, mesh_tables, tri_ids_of
from emerge_b200 import bc as B
a, b, L = 22.86e-3, 10.16e-3, 45e-3
box = box_mesh(6, 3, 12, a, b, L, jitter=0.1, seed=2)
t = mesh_tables(box.nodes_xyz, box.tets)
nT = t.tets.shape[1]
er = np.repeat(np.eye(3, dtype=complex)[:, :, None], nT, axis=2)
ur = er.copy()
tag = lambda k: tri_ids_of(t, box.face_tris[box.face_tag == k])
bcs = [B.PEC(np.concatenate([tag(k) for k in (1, 2, 3, 4)])),
       B.RectangularWaveguide(tag(5), 1, B.CoordSys(origin=(0, 0, 0)), (a, b)),
       B.RectangularWaveguide(tag(6), 2, B.CoordSys(origin=(0, 0, L)), (a, b))]
freqs = np.linspace(8e9, 12e9, 41)
for rtol in (1e-9, 1e-8):
    warm = FrequencySweep(t, er, ur, bcs, recycle=32)
    warm.solver_opts.update(rtol=rtol)
    rw = warm.run(freqs)
    info = warm.ctx.recycle_info()
    warm.ctx.close()
    its = [s["iters"] for s in rw.stats[::2]]
    print(f"rtol={rtol}: free points {sum(1 for i in its if i == 0)}/41, iterations {sum(its)}, "
          f"directions {info['n']}, max relres {max(s['relres'] for s in rw.stats):.2e}", flush=True)
    print("   iters per point (sweep order of the frequencies):", its, flush=True)
