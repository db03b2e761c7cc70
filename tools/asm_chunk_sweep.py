"""Numeric-phase timing against the chunk size of the element-kernel -> reduction pipeline (L2-resident COO scratch).
usage: asm_chunk_sweep.py nx,ny,nz   -> prints one JSON line per configuration; checks bitwise equality with the single pass"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from emerge_b200.lib import Context  # noqa: E402

cells = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "24,12,80").split(","))
box, t, er, ur, bcs, L = bench.make_waveguide(*cells)
nT = t.tets.shape[1]
ctx = Context(0)
ctx.upload_mesh(t.nodes, t.tets, t.tris, t.tet_to_field, t.tri_to_field, t.edges.shape[1])
ctx.upload_materials(er, ur)
ctx.symbolic()
ref = None
out = []
for chunk, persist in [(nT, 0), (2368, 1), (4736, 1), (4736, 0), (9472, 1), (9472, 0), (18944, 1), (37888, 1), (37888, 0)]:
    ctx.assemble_config(chunk, bool(persist))
    ms = []
    for rep in range(3):
        ctx.assemble_KM()
        ms.append(ctx.last_ms("assemble"))
    _, _, K = ctx.get_csr(0, pattern=False)
    _, _, M = ctx.get_csr(1, pattern=False)
    if ref is None:
        ref = (K.copy(), M.copy())
    same = bool(np.array_equal(K.view(np.float64), ref[0].view(np.float64)) and np.array_equal(M.view(np.float64), ref[1].view(np.float64)))
    rec = dict(tets=nT, chunk=chunk, persist=persist, ms=ms, Mtet_per_s=nT / (min(ms) * 1e3), bitwise_equal_to_single_pass=same)
    print(json.dumps(rec), flush=True)
    out.append(rec)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/asm_chunk_sweep.json", "w"), indent=1)
