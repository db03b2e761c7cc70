"""Numeric-phase timing of both assembly algorithms (development probe / profiles).
usage: asm_bench.py nx,ny,nz [reps]   (EMB_PROFILE=1: open a cudaProfilerStart/Stop window around one fused assembly)"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from emerge_b200.lib import Context  # noqa: E402

cells = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "44,20,190").split(","))
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
box, t, er, ur, bcs, L = bench.make_waveguide(*cells)
ctx = Context(0)
ctx.upload_mesh(t.nodes, t.tets, t.tris, t.tet_to_field, t.tri_to_field, t.edges.shape[1])
ctx.upload_materials(er, ur)
ctx.symbolic()
nT = t.tets.shape[1]
out = {"tets": int(nT), "nnz": int(ctx.lib.emb_nnz(ctx.h)), "symbolic_ms": ctx.last_ms("symbolic"),
       "symbolic_entities_ms": ctx.last_ms("symbolic_entities")}
vals = {}
for mode in ("fused", "coo"):
    ctx.assemble_mode(mode)
    ts = []
    for r in range(reps):
        if mode == "fused" and r == reps - 1 and os.environ.get("EMB_PROFILE") == "1":
            ctx.profiler(True)
        ctx.assemble_KM()
        if mode == "fused" and r == reps - 1 and os.environ.get("EMB_PROFILE") == "1":
            ctx.profiler(False)
        ts.append((ctx.last_ms("tet_kernel"), ctx.last_ms("reduce")))
    a, b = min(ts, key=lambda v: v[0] + v[1])
    out[mode] = {"first_kernel_ms": a, "second_kernel_ms": b, "total_ms": a + b, "Mtet_per_s": nT / (a + b) / 1e3,
                 "algorithmic_GBps": 9722 * nT / (a + b) / 1e6, "all": ts}
    vals[mode] = (ctx.get_csr(0, pattern=False)[2], ctx.get_csr(1, pattern=False)[2])
out["fused_vs_coo_rel_diff"] = [float(np.abs(a - b).max() / np.abs(b).max()) for a, b in zip(vals["fused"], vals["coo"])]
print(json.dumps(out))
