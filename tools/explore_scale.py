"""GPU exploration: timings of every phase on a synthetic waveguide of given size (scratch tool)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from emerge_b200.synthmesh import box_mesh, mesh_tables, tri_ids_of
from emerge_b200 import bc as B
from emerge_b200.sweep import FrequencySweep

def make(nx, ny, nz, jitter=0.0):
    a, b = 22.86e-3, 10.16e-3
    L = nz * a / nx
    t0 = time.time()
    box = box_mesh(nx, ny, nz, a, b, L, jitter=jitter)
    t = mesh_tables(box.nodes_xyz, box.tets)
    nT = t.tets.shape[1]
    er = np.zeros((3, 3, nT), complex); er[0, 0] = er[1, 1] = er[2, 2] = 1
    ur = er.copy()
    tag = lambda k: tri_ids_of(t, box.face_tris[box.face_tag == k])
    bcs = [B.PEC(np.concatenate([tag(k) for k in (1, 2, 3, 4)])),
           B.RectangularWaveguide(tag(5), 1, B.CoordSys(origin=(0, 0, 0)), (a, b)),
           B.RectangularWaveguide(tag(6), 2, B.CoordSys(origin=(0, 0, L)), (a, b))]
    print(f"mesh {nx}x{ny}x{nz}: nT={nT} N={t.n_field} host mesh+tables {time.time()-t0:.1f}s", flush=True)
    return t, er, ur, bcs, (a, b, L)

if __name__ == "__main__":
    nx, ny, nz = [int(v) for v in sys.argv[1:4]]
    maxit = int(sys.argv[4]) if len(sys.argv) > 4 else 2000
    precond = sys.argv[5] if len(sys.argv) > 5 else "block"
    t, er, ur, bcs, dims = make(nx, ny, nz)
    sw = FrequencySweep(t, er, ur, bcs)
    t0 = time.time(); sw.setup(); print("setup wall", time.time() - t0, sw.timings, flush=True)
    ctx = sw.ctx
    print("nnz", ctx.nnz, "Ns", ctx.n_solve, "nnz_s", ctx.lib.emb_csr_nnz(ctx.h, 2))
    for rep in range(2):
        ctx.assemble_KM(); print("assemble: tet", ctx.last_ms("tet_kernel"), "reduce", ctx.last_ms("reduce"))
    k0 = sw.assemble_frequency(10e9); print("form_A ms", ctx.last_ms("form_A"))
    ms = ctx.spmv_bench(20); nnz_s = ctx.lib.emb_csr_nnz(ctx.h, 2); Ns = ctx.n_solve
    byt = 20 * nnz_s + 36 * Ns + 4
    print(f"spmv {ms:.4f} ms  {byt/ms/1e6:.1f} GB/s algorithmic")
    for pc in precond.split(","):
        x, info = ctx.solve(0, want_x=False, raise_on_fail=False, method="cocr", precond=pc, rtol=1e-8, maxit=maxit)
        print("solve", pc, info, flush=True)
