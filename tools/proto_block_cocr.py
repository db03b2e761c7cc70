"""Prototype (CPU, scratch): block COCR (one Krylov space for all ports) against independent lockstep recurrences inside the
defect-correction loop.  usage: proto_block_cocr.py nx ny nz   (20k tets: 544 -> 351 inner iterations for two ports)"""
import sys
sys.argv = [sys.argv[0]] + sys.argv[1:4] + ["none"]
import os
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import numpy as np
exec(open(os.path.join(HERE, 'proto_hx.py')).read().split("for v in variants:")[0])
Minv1 = prec("add")
Minv = lambda R: np.stack([Minv1(R[:, k]) for k in range(R.shape[1])], axis=1)
def bcocr_inner(RHS, stop, maxit=3000, block=True):
    n, s = RHS.shape
    X = np.zeros_like(RHS); R = RHS.copy(); Z = Minv(R); P = Z.copy(); AZ = As @ Z; AP = AZ.copy()
    rho = Z.T @ AZ if block else np.diag(np.einsum('ik,ik->k', Z, AZ))
    for it in range(1, maxit + 1):
        MAP = Minv(AP)
        sig = AP.T @ MAP if block else np.diag(np.einsum('ik,ik->k', AP, MAP))
        alpha = np.linalg.solve(sig, rho)
        X = X + P @ alpha; R = R - AP @ alpha; Z = Z - MAP @ alpha
        rn = np.linalg.norm(R, axis=0)
        if np.all(rn <= stop): break
        AZ = As @ Z
        rho_new = Z.T @ AZ if block else np.diag(np.einsum('ik,ik->k', Z, AZ))
        beta = np.linalg.solve(rho, rho_new); rho = rho_new
        P = Z + P @ beta; AP = AZ + AP @ beta
    return X, it
rng = np.random.default_rng(0)
for name, B, block in (("lockstep s=2", np.stack(rhs, axis=1), False), ("block s=2", np.stack(rhs, axis=1), True),
                       ("block s=4 (2 ports + 2 random)", np.stack(rhs + [rng.standard_normal(len(rhs[0])) + 0j, rng.standard_normal(len(rhs[0])) + 0j], axis=1), True)):
    bn = np.linalg.norm(B, axis=0)
    X = np.zeros_like(B); tot = 0; log = []
    for outer in range(40):
        R = B - A @ X; rn = np.linalg.norm(R, axis=0)
        log.append(f"{(rn/bn)[:2].max():.1e}")
        if np.all((rn / bn)[:2] <= 1e-9): break
        stop = np.maximum(1e-2 * rn, 0.3e-9 * bn)
        if B.shape[1] > 2: stop[2:] = 1e300      # the extra columns only feed the Krylov space
        D, it = bcocr_inner(R, stop, block=block); tot += it; log.append(f"[{it}]")
        X = X + D
    print(name, "total inner iterations", tot, "outer", outer, " ".join(log), flush=True)
