"""Tuning sweep of the staged SELL operator application (k_bsell_tma2<Q, WARPS, META, HINT>, sell.cuh): ms per launch and a
checksum of y for a fixed pseudo-random x (all variants add in the same order: the checksums must be identical).
usage: sell_variants.py nx,ny,nz [variant ...]     variant = Q:WARPS:DEPTH:HINT, "default" = the shipped configuration"""
import os
import sys

os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
os.environ["EMB_SPMV_CHECK"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from emerge_b200.sweep import FrequencySweep  # noqa: E402

cells = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "24,12,80").split(","))
variants = sys.argv[2:] or ["default", "8:24:2:1", "8:24:2:0", "8:24:3:1", "8:24:2:5", "8:24:2:3", "6:28:2:1", "5:32:2:1", "10:20:2:1",
                            "16:16:2:1", "16:16:2:0", "default"]
box, t, er, ur, bcs, L = bench.make_waveguide(*cells)
sw = FrequencySweep(t, er, ur, bcs, device=0, recycle=0)
sw.solver_opts.update(precond="block")
sw.setup()
ctx = sw.ctx
sw.assemble_frequency(10e9)
nnz, Ns = int(ctx.lib.emb_csr_nnz(ctx.h, 2)), ctx.n_solve
b = 9 * nnz + (4 + 64) * Ns
ref = None
for v in variants:
    if v == "default":
        os.environ.pop("EMB_SELL_VARIANT", None)
    else:
        os.environ["EMB_SELL_VARIANT"] = v
    ms = ctx.spmv_bench(20, nv=2, fp32=True)
    cs = tuple(ctx.last_ms(k) for k in ("spmv_check_re", "spmv_check_im", "spmv_check_abs2"))
    if ref is None:
        ref = cs
    print(f"variant {v:>10}: {ms:.4f} ms  {b / ms / 1e6:.0f} GB/s  checksum {'== first' if cs == ref else 'DIFFERS ' + repr(cs) + ' vs ' + repr(ref)}"
          f"  |y - A x|/|A x| {ctx.last_ms('spmv_check_vs_A'):.2e}", flush=True)
