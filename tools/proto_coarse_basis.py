"""Prototype (CPU, scratch): the reduced basis U (solutions at other frequencies) as an extra coarse space of the
preconditioner,  M^-1 += U (U^T As U)^-1 U^T,  on top of the minimum-residual start vector.
usage: proto_coarse_basis.py nx ny nz"""
import sys
import os
HERE = os.path.dirname(os.path.abspath(__file__))
args = sys.argv[1:4]
sys.path.insert(0, HERE)
import numpy as np
SRC = open(os.path.join(HERE, 'proto_hx.py')).read().split("for v in variants:")[0]

def setup(fval):
    ns = {"__name__": "proto"}
    sys.argv = [sys.argv[0]] + args + ["none"]
    exec(SRC.replace("f = 10e9", f"f = {fval!r}"), ns)
    return ns

def bcocr(As, Minv, RHS, stop, X0=None, maxit=4000):
    X = np.zeros_like(RHS) if X0 is None else X0.copy()
    R = RHS - (As @ X if X0 is not None else 0)
    Z = Minv(R); P = Z.copy(); AZ = As @ Z; AP = AZ.copy(); rho = Z.T @ AZ
    for it in range(1, maxit + 1):
        MAP = Minv(AP); sig = AP.T @ MAP
        alpha = np.linalg.solve(sig, rho)
        X = X + P @ alpha; R = R - AP @ alpha; Z = Z - MAP @ alpha
        if np.all(np.linalg.norm(R, axis=0) <= stop): break
        AZ = As @ Z; rho_new = Z.T @ AZ; beta = np.linalg.solve(rho, rho_new); rho = rho_new
        P = Z + P @ beta; AP = AZ + AP @ beta
    return X, it

def solve(ns, U=None, coarse=False, tol=3e-9):
    A, As, rhs = ns["A"], ns["As"], ns["rhs"]
    M1 = ns["prec"]("add")
    Minv = lambda R: np.stack([M1(R[:, k]) for k in range(R.shape[1])], axis=1)
    if coarse and U is not None:
        G = U.T @ (As @ U)
        base = Minv
        Minv = lambda R: base(R) + U @ np.linalg.solve(G, U.T @ R)
    B = np.stack(rhs, axis=1); bn = np.linalg.norm(B, axis=0)
    X = np.zeros_like(B)
    if U is not None:
        C = A @ U
        y, *_ = np.linalg.lstsq(C, B, rcond=None)
        X = U @ y
    tot = 0; log = []
    for outer in range(40):
        R = B - A @ X; rn = np.linalg.norm(R, axis=0)
        log.append(f"{(rn / bn).max():.1e}")
        if np.all(rn / bn <= tol): break
        D, it = bcocr(As, Minv, R, np.maximum(1e-2 * rn, 0.3 * tol * bn)); tot += it; log.append(f"[{it}]")
        X = X + D
    return X, tot, " ".join(log)

nsA, nsB, nsC = setup(9.0e9), setup(11.0e9), setup(10.0e9)
XA, itA, lA = solve(nsA); print("9 GHz cold      ", itA, lA, flush=True)
XB, itB, lB = solve(nsB); print("11 GHz cold     ", itB, lB, flush=True)
U = np.concatenate([XA, XB], axis=1)
U = U / np.linalg.norm(U, axis=0)
_, it0, l0 = solve(nsC); print("10 GHz cold     ", it0, l0, flush=True)
_, it1, l1 = solve(nsC, U); print("10 GHz projected", it1, l1, flush=True)
_, it2, l2 = solve(nsC, U, coarse=True); print("10 GHz proj+coarse", it2, l2, flush=True)
