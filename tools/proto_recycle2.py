"""Prototype (CPU, scratch): recycling with noisy snapshots (iterative-solver accuracy) in sequential vs hierarchical order.
Prints, per point, the projected relres; counts points needing a Krylov solve (relres > rtol)."""
import sys, time
import numpy as np, scipy.sparse.linalg as spla
from proto_common import *

nx, ny, nz = [int(v) for v in sys.argv[1:4]]
nf = int(sys.argv[4]); order = sys.argv[5]; tol = float(sys.argv[6]); cap = int(sys.argv[7])
rtol = 1e-8
L = nz * 22.86e-3 / nx
S = waveguide_system(nx, ny, nz, L)
freqs = np.linspace(8e9, 12e9, nf)
def hier(lo, hi):
    out = [lo, hi]; seg = [(lo, hi)]
    while seg:
        nxt = []
        for a, b in seg:
            if b - a > 1:
                m = (a + b) // 2; out.append(m); nxt += [(a, m), (m, b)]
        seg = nxt
    return out
idx = hier(0, nf - 1) if order == "hier" else list(range(nf))
rng = np.random.default_rng(0)
U = []   # newest first
cold = 0
for n, i in enumerate(idx):
    f = freqs[i]
    A, rhs = system_at(S, f)
    lu = spla.splu(A.tocsc())
    if U:
        W = A @ np.array(U).T
        Q, R = np.linalg.qr(W)
    msg = []
    for r in rhs:
        x0 = np.zeros_like(r)
        if U:
            y = np.linalg.solve(R, Q.conj().T @ r)
            x0 = np.array(U).T @ y
        rr = np.linalg.norm(r - A @ x0) / np.linalg.norm(r)
        msg.append(rr)
        if rr > rtol:
            cold += 1
            noise = rng.standard_normal(len(r)) + 1j * rng.standard_normal(len(r))
            noise *= tol * np.linalg.norm(r) / np.linalg.norm(noise)
            x = lu.solve(r + noise)
            d = x - x0
            U.insert(0, d / np.linalg.norm(d))
            if len(U) > cap: U.pop()
            W = A @ np.array(U).T; Q, R = np.linalg.qr(W)
    print(f"{n:3d} i={i:3d} f={f/1e9:.3f} nU={len(U):3d} proj " + " ".join(f"{v:.1e}" for v in msg), flush=True)
print("krylov solves needed:", cold, "of", 2 * nf)
