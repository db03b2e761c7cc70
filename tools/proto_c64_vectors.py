"""Does the inner COCR survive complex64 vectors?  Emulation: defect correction on the true A, inner COCR on As with every
vector rounded to complex64 after each update (dots in float64), inner reduction 1e-2 per outer step."""
import sys
sys.argv = [sys.argv[0]] + sys.argv[1:4] + ["none"]
import os
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import numpy as np
exec(open(os.path.join(HERE, 'proto_hx.py')).read().split("for v in variants:")[0])
Minv = prec("add")
c64 = lambda v: v.astype(np.complex64).astype(np.complex128)
def cocr_inner(rhs, stop, lowp, maxit=3000):
    rd = c64 if lowp else (lambda v: v)
    x = np.zeros_like(rhs); r = rd(rhs.copy()); z = rd(Minv(r)); p = z.copy(); Az = rd(As @ z); Ap = Az.copy(); zAz = z @ Az
    for it in range(1, maxit + 1):
        MAp = rd(Minv(Ap)); alpha = zAz / (Ap @ MAp)
        x = rd(x + alpha * p); r = rd(r - alpha * Ap); z = rd(z - alpha * MAp)
        if np.linalg.norm(r) <= stop: break
        Az = rd(As @ z); zn = z @ Az; beta = zn / zAz; zAz = zn
        p = rd(z + beta * p); Ap = rd(Az + beta * Ap)
    return x, it
b = rhs[0]; bn = np.linalg.norm(b)
for lowp in (False, True):
    x = np.zeros_like(b); tot = 0; log = []
    for outer in range(40):
        r = b - A @ x; rn = np.linalg.norm(r)
        log.append(f"{rn/bn:.1e}")
        if rn / bn <= 1e-9: break
        d, it = cocr_inner(r, max(1e-2 * rn, 0.3e-9 * bn), lowp); tot += it
        x = x + d
    print(("complex64 vectors" if lowp else "complex128 vectors"), "total inner iterations", tot, "outer", outer, " ".join(log), flush=True)
