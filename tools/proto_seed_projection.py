"""Prototype (CPU, scratch): plateau-free restarts of the defect correction.  The block search directions of the inner
solves of one frequency point are kept (P_i, M^-1 A P_i, sigma_i); every later inner solve starts from the projection of
its right-hand side onto that space (seed projection in the M^-1 inner product the block COCR is conjugate in).
usage: proto_seed_projection.py nx ny nz [max stored block iterations]"""
import sys
import os
HERE = os.path.dirname(os.path.abspath(__file__))
keep_max = int(sys.argv[4]) if len(sys.argv) > 4 else 10 ** 9
sys.argv = [sys.argv[0]] + sys.argv[1:4] + ["none"]
sys.path.insert(0, HERE)
import numpy as np
exec(open(os.path.join(HERE, 'proto_hx.py')).read().split("for v in variants:")[0])
Minv1 = prec("add")
Minv = lambda R: np.stack([Minv1(R[:, k]) for k in range(R.shape[1])], axis=1)

def bcocr(RHS, stop, store, X0=None, maxit=3000):
    X = np.zeros_like(RHS) if X0 is None else X0.copy()
    R = RHS - (As @ X if X0 is not None else 0)
    Z = Minv(R); P = Z.copy(); AZ = As @ Z; AP = AZ.copy(); rho = Z.T @ AZ
    for it in range(1, maxit + 1):
        MAP = Minv(AP); sig = AP.T @ MAP
        if store is not None and len(store) < keep_max: store.append((P.copy(), MAP.copy(), sig.copy()))
        alpha = np.linalg.solve(sig, rho)
        X = X + P @ alpha; R = R - AP @ alpha; Z = Z - MAP @ alpha
        if np.all(np.linalg.norm(R, axis=0) <= stop): break
        AZ = As @ Z; rho_new = Z.T @ AZ; beta = np.linalg.solve(rho, rho_new); rho = rho_new
        P = Z + P @ beta; AP = AZ + AP @ beta
    return X, it

B = np.stack(rhs, axis=1); bn = np.linalg.norm(B, axis=0)
for mode in ("restart from zero", "seed projection"):
    X = np.zeros_like(B); tot = 0; log = []; store = [] if mode != "restart from zero" else None
    for outer in range(40):
        R = B - A @ X; rn = np.linalg.norm(R, axis=0)
        log.append(f"{(rn / bn).max():.1e}")
        if np.all(rn / bn <= 3e-9): break
        stop = np.maximum(1e-2 * rn, 0.3 * 3e-9 * bn)
        X0 = None
        if store:
            X0 = np.zeros_like(R)
            for P_i, MAP_i, sig_i in store:
                X0 = X0 + P_i @ np.linalg.solve(sig_i, MAP_i.T @ R)
            log.append(f"(proj {np.linalg.norm(R - As @ X0, axis=0).max() / rn.max():.1e})")
        D, it = bcocr(R, stop, store, X0); tot += it; log.append(f"[{it}]")
        X = X + D
    print(f"{mode:18s} total inner iterations {tot:4d} outer {outer}  " + " ".join(log), flush=True)
