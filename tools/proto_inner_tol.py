"""Prototype (CPU, scratch): total inner iterations of the defect correction against the inner tolerance.
usage: proto_inner_tol.py nx ny nz"""
import sys
sys.argv = [sys.argv[0]] + sys.argv[1:4] + ["none"]
import os
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import numpy as np
exec(open(os.path.join(HERE, 'proto_hx.py')).read().split("for v in variants:")[0])
Minv = prec("add")
def cocr_inner(rhs, stop, maxit=3000):
    x = np.zeros_like(rhs); r = rhs.copy(); z = Minv(r); p = z.copy(); Az = As @ z; Ap = Az.copy(); zAz = z @ Az
    for it in range(1, maxit + 1):
        MAp = Minv(Ap); alpha = zAz / (Ap @ MAp)
        x = x + alpha * p; r = r - alpha * Ap; z = z - alpha * MAp
        if np.linalg.norm(r) <= stop: break
        Az = As @ z; zn = z @ Az; beta = zn / zAz; zAz = zn
        p = z + beta * p; Ap = Az + beta * Ap
    return x, it
b = rhs[0]; bn = np.linalg.norm(b)
for red in (1e-2, 3e-3, 1e-3, 1e-4):
    x = np.zeros_like(b); tot = 0; log = []
    for outer in range(40):
        r = b - A @ x; rn = np.linalg.norm(r)
        log.append(f"{rn/bn:.1e}")
        if rn / bn <= 1e-9: break
        d, it = cocr_inner(r, max(red * rn, 0.3e-9 * bn)); tot += it; log.append(f"[{it}]")
        x = x + d
    print("inner_red", red, "total inner iterations", tot, "outer", outer, " ".join(log), flush=True)
