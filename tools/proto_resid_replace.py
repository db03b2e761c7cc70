"""Prototype (CPU, scratch): ONE continuous COCR run on the symmetric part As with the recurrence residual REPLACED by the
true residual b - A x of the non-symmetric operator whenever it has dropped by `fac` since the last replacement, against the
restarted defect correction of the product.  usage: proto_resid_replace.py nx ny nz"""
import sys
sys.argv = [sys.argv[0]] + sys.argv[1:4] + ["none"]
import os
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import numpy as np
exec(open(os.path.join(HERE, 'proto_hx.py')).read().split("for v in variants:")[0])
Minv = prec("add")
b = rhs[0]; bn = np.linalg.norm(b)

def cocr_inner(rhs_, stop, maxit=3000):
    x = np.zeros_like(rhs_); r = rhs_.copy(); z = Minv(r); p = z.copy(); Az = As @ z; Ap = Az.copy(); zAz = z @ Az
    for it in range(1, maxit + 1):
        MAp = Minv(Ap); alpha = zAz / (Ap @ MAp)
        x = x + alpha * p; r = r - alpha * Ap; z = z - alpha * MAp
        if np.linalg.norm(r) <= stop: break
        Az = As @ z; zn = z @ Az; beta = zn / zAz; zAz = zn
        p = z + beta * p; Ap = Az + beta * Ap
    return x, it

def restarted(red=1e-2, rtol=1e-9):
    x = np.zeros_like(b); tot = 0
    for outer in range(40):
        r = b - A @ x; rn = np.linalg.norm(r)
        if rn / bn <= rtol: break
        d, it = cocr_inner(r, max(red * rn, 0.3 * rtol * bn)); tot += it
        x = x + d
    return tot, rn / bn

def replaced(fac, rtol=1e-9, maxit=3000, keep_dir=True):
    """extra work per replacement: one A x, one M^-1, (one As z is done anyway)"""
    x = np.zeros_like(b); r = b.copy(); z = Minv(r); p = z.copy(); Az = As @ z; Ap = Az.copy(); zAz = z @ Az
    last = bn; nrep = 0; log = []
    for it in range(1, maxit + 1):
        MAp = Minv(Ap); alpha = zAz / (Ap @ MAp)
        x = x + alpha * p; r = r - alpha * Ap; z = z - alpha * MAp
        rn = np.linalg.norm(r)
        if rn <= fac * last or rn <= 0.5 * rtol * bn:
            rt = b - A @ x; rtn = np.linalg.norm(rt); nrep += 1
            log.append(f"it {it}: recurrence {rn/bn:.1e} true {rtn/bn:.1e}")
            if rtn <= rtol * bn:
                return it, nrep, rtn / bn, log
            r = rt; z = Minv(r); last = rtn
            if not keep_dir:
                p = z.copy(); Az = As @ z; Ap = Az.copy(); zAz = z @ Az
                continue
        Az = As @ z; zn = z @ Az; beta = zn / zAz; zAz = zn
        p = z + beta * p; Ap = Az + beta * Ap
    return it, nrep, np.linalg.norm(b - A @ x) / bn, log

print("restarted defect correction (product):", restarted(), flush=True)
for fac in (0.3, 0.1, 0.03, 0.01):
    it, nrep, rr, log = replaced(fac)
    print(f"replacement at drops of {fac}: {it} iterations, {nrep} replacements, true relres {rr:.2e}", flush=True)
    print("    " + "; ".join(log[:12]), flush=True)
