"""Prototype (CPU, scratch): how well do previous solutions of the sweep predict the next frequency point?
For each new frequency the min-residual combination of the stored snapshots (both ports) is computed and the
relative residual ||b - A(f) V y|| / ||b|| it leaves is printed."""
import sys, time
import numpy as np, scipy.sparse.linalg as spla
from proto_common import *

nx, ny, nz = [int(v) for v in sys.argv[1:4]]
nf = int(sys.argv[4]) if len(sys.argv) > 4 else 41
order = sys.argv[5] if len(sys.argv) > 5 else "seq"
L = nz * 22.86e-3 / nx
S = waveguide_system(nx, ny, nz, L)
print("N", S['N'], "solve", len(S['solve_ids']), flush=True)
freqs = np.linspace(8e9, 12e9, nf)
if order == "bisect":
    idx = [0, nf - 1]
    step = nf - 1
    while step > 1:
        half = step // 2
        idx += [i for i in range(half, nf - 1, step) if i not in idx]
        step = half
    idx += [i for i in range(nf) if i not in idx]
else:
    idx = list(range(nf))
V = []
for n, i in enumerate(idx):
    f = freqs[i]
    A, rhs = system_at(S, f)
    t0 = time.time()
    lu = spla.splu(A.tocsc())
    xs = [lu.solve(r) for r in rhs]
    res = []
    if V:
        Vm = np.array(V).T
        W = A @ Vm
        Q, R = np.linalg.qr(W)
        for r in rhs:
            y = Q.conj().T @ r
            res.append(np.linalg.norm(r - Q @ y) / np.linalg.norm(r))
    print(f"{n:3d} f={f/1e9:.3f} nV={len(V):3d} proj relres " + " ".join(f"{v:.2e}" for v in res) + f"  ({time.time()-t0:.1f}s)", flush=True)
    V.extend(xs)
