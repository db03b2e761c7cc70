"""Prototype (CPU, scratch): the corrections of the earlier outer steps of the SAME point as a coarse space of the
preconditioner of the later ones, M^-1 += D (D^T As D)^-1 D^T.  usage: proto_outer_coarse.py nx ny nz"""
import sys
sys.argv = [sys.argv[0]] + sys.argv[1:4] + ["none"]
import os
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import numpy as np
exec(open(os.path.join(HERE, 'proto_hx.py')).read().split("for v in variants:")[0])
M0 = prec("add")
b = rhs[0]; bn = np.linalg.norm(b)

def cocr_inner(rhs_, stop, Minv, maxit=3000, keep=0, keep_p=False):
    x = np.zeros_like(rhs_); r = rhs_.copy(); z = Minv(r); p = z.copy(); Az = As @ z; Ap = Az.copy(); zAz = z @ Az
    snaps = []
    for it in range(1, maxit + 1):
        MAp = Minv(Ap); alpha = zAz / (Ap @ MAp)
        x = x + alpha * p; r = r - alpha * Ap; z = z - alpha * MAp
        if keep and it % keep == 0: snaps.append(x.copy() if not keep_p else p.copy())
        if np.linalg.norm(r) <= stop: break
        Az = As @ z; zn = z @ Az; beta = zn / zAz; zAz = zn
        p = z + beta * p; Ap = Az + beta * Ap
    return x, it, snaps

def run(mode, red=1e-2, rtol=1e-9, keep=0, first_only=False, last=0, ritz=0, diffs=False):
    x = np.zeros_like(b); tot = 0; D = []; log = []
    for outer in range(40):
        r = b - A @ x; rn = np.linalg.norm(r)
        if rn / bn <= rtol: break
        if D and mode != "none":
            Dm = np.stack(D[-last:] if last else D, axis=1)
            if diffs:
                Dm = np.concatenate([Dm[:, :1], np.diff(Dm, axis=1)], axis=1)
            Q, _ = np.linalg.qr(Dm)
            if ritz and Q.shape[1] > ritz:
                Gq = Q.T @ (As @ Q)
                _, sv, Vh = np.linalg.svd(Gq)
                Q = Q @ Vh.conj().T[:, -ritz:]          # right singular vectors of the smallest singular values
                Q, _ = np.linalg.qr(Q)
            G = Q.T @ (As @ Q)
            Gi = np.linalg.inv(G)
            Minv = lambda v, Q=Q, Gi=Gi: M0(v) + Q @ (Gi @ (Q.T @ v))
        else:
            Minv = M0
        d, it, snaps = cocr_inner(r, max(red * rn, 0.3 * rtol * bn), Minv, keep=keep, keep_p=(mode == 'pdirs')); tot += it; log.append(it)
        x = x + d
        if mode == "corr": D.append(d)
        if mode in ("snaps", "pdirs") and (outer == 0 or not first_only): D.extend(snaps + [d])
    return tot, log, rn / bn

print("plain restarts:", run("none"), flush=True)
print("every 10, all 14:", run("snaps", keep=10, first_only=True), flush=True)
print("every 10, last 8:", run("snaps", keep=10, first_only=True, last=8), flush=True)
print("every 5, last 8:", run("snaps", keep=5, first_only=True, last=8), flush=True)
print("every 5, last 12:", run("snaps", keep=5, first_only=True, last=12), flush=True)
print("every 10, 8 smallest-singular-value combinations of all:", run("snaps", keep=10, first_only=True, ritz=8), flush=True)
print("every 10, 4 smallest-singular-value combinations of all:", run("snaps", keep=10, first_only=True, ritz=4), flush=True)
print("every 5, 8 smallest-singular-value combinations of all:", run("snaps", keep=5, first_only=True, ritz=8), flush=True)
