"""Prototype (CPU, scratch): multilevel auxiliary-space preconditioner variants for A(f), iteration counts of COCR."""
import sys, time
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
from proto_common import *
from emerge_b200.auxspace import build_aux_spaces
from emerge_b200.amg import sa_hierarchy

nx, ny, nz = [int(v) for v in sys.argv[1:4]]
variants = sys.argv[4].split(",") if len(sys.argv) > 4 else ["jac", "mg"]
f = 10e9
L = nz * 22.86e-3 / nx
S = waveguide_system(nx, ny, nz, L)
t = S['t']
A, rhs = system_at(S, f)
A = A.tocsr()
As = ((A + A.T) * 0.5).tocsr()
k0 = 2 * np.pi * f / 299792458
sid = S['solve_ids']
N = S['N']
ns = len(sid)
print(f"tets {t.tets.shape[1]} N {N} ns {ns} nnz {A.nnz}", flush=True)
Ks = S['E'].tocsr()[sid][:, sid].real.tocsr()
Ms = S['B'].tocsr()[sid][:, sid].real.tocsr()
Ms = ((Ms + Ms.T) * 0.5).tocsr()

t0 = time.time()
G, P, G1 = build_aux_spaces(t)
nN, nE = t.nodes.shape[1], t.edges.shape[1]
keep = np.zeros(N, bool); keep[sid] = True
elim = ~keep
def restrict(R):
    bad = np.asarray(abs(R[elim]).sum(axis=0)).ravel() > 0
    return R[keep][:, ~bad].tocsr(), bad
Gs, badG = restrict(G)
Ps, badP = restrict(P)
# P1 -> P2 injection on kept dofs
e = np.asarray(t.edges)
I21 = sp.coo_matrix((np.concatenate([np.ones(nN), 0.5 * np.ones(nE), 0.5 * np.ones(nE)]),
                     (np.concatenate([np.arange(nN), nN + np.arange(nE), nN + np.arange(nE)]),
                      np.concatenate([np.arange(nN), e[0], e[1]]))), shape=(nN + nE, nN)).tocsr()
badN = np.asarray(abs(I21[badG]).sum(axis=0)).ravel() > 0      # P1 nodes touching dropped P2 dofs
I21s = I21[~badG][:, ~badN].tocsr()
# nodal vector interpolation Pi (nE x 3nN), node-major interleaved
d = (t.nodes[:, e[1]] - t.nodes[:, e[0]])                       # (3,nE)
rows = np.repeat(np.arange(nE), 6)
cols = np.stack([3 * e[0] + 0, 3 * e[0] + 1, 3 * e[0] + 2, 3 * e[1] + 0, 3 * e[1] + 1, 3 * e[1] + 2], axis=1).ravel()
vals = np.stack([-0.5 * d[0], -0.5 * d[1], -0.5 * d[2], -0.5 * d[0], -0.5 * d[1], -0.5 * d[2]], axis=1).ravel()
Pi = sp.coo_matrix((vals, (rows, cols)), shape=(nE, 3 * nN)).tocsr()
badV = np.asarray(abs(Pi[badP]).sum(axis=0)).ravel() > 0
badNode = badV.reshape(nN, 3).any(axis=1)
keepV = np.repeat(~badNode, 3)
Pis = Pi[~badP][:, keepV].tocsr()
print(f"aux spaces {time.time()-t0:.1f}s  G {Gs.shape} P {Ps.shape} I21 {I21s.shape} Pi {Pis.shape}", flush=True)

t0 = time.time()
L2 = (Gs.T @ Ms @ Gs).tocsr()
L1 = (I21s.T @ L2 @ I21s).tocsr()
Kw = (Ps.T @ Ks @ Ps).tocsr(); Mw = (Ps.T @ Ms @ Ps).tocsr()
tau = k0 ** 2
An = (Pis.T @ (Kw + tau * Mw) @ Pis).tocsr()
print(f"galerkin {time.time()-t0:.1f}s L2 {L2.shape} nnz/row {L2.nnz/L2.shape[0]:.1f} L1 nnz/row {L1.nnz/L1.shape[0]:.1f} An {An.shape} nnz/row {An.nnz/An.shape[0]:.1f}", flush=True)
t0 = time.time()
H1 = sa_hierarchy(L1, block=1)
Hn = sa_hierarchy(An, block=3)
print(f"amg setup {time.time()-t0:.1f}s  L1 levels {[l['A'].shape[0] for l in H1]}  An levels {[l['A'].shape[0] for l in Hn]}", flush=True)
print("rho", [round(l['rho'], 2) for l in H1], [round(l['rho'], 2) for l in Hn])

def vcycle(H, b, lvl=0, nu=1, w=0.7):
    lev = H[lvl]
    Al = lev['A']
    if lev['P'] is None:
        if 'lu' not in lev: lev['lu'] = spla.splu(sp.csc_matrix(Al + 1e-12 * sp.identity(Al.shape[0])))
        return lev['lu'].solve(b.real) + 1j * lev['lu'].solve(b.imag)
    om = w * 2.0 / lev['rho'] if False else 1.0 / lev['rho'] * 4.0 / 3.0 * w / 0.7 * 0.7
    om = 4.0 / (3.0 * lev['rho'])
    x = om * lev['dinv'] * b
    for _ in range(nu - 1): x = x + om * lev['dinv'] * (b - Al @ x)
    rc = lev['P'].T @ (b - Al @ x)
    x = x + lev['P'] @ vcycle(H, rc, lvl + 1, nu)
    for _ in range(nu): x = x + om * lev['dinv'] * (b - Al @ x)
    return x

# fine-level block Jacobi on As (2x2 edge/face pairs)
nTri = t.tris.shape[1]
newid = -np.ones(N, int); newid[sid] = np.arange(ns)
mate_full = np.where(np.arange(N) < nE + nTri, np.arange(N) + nE + nTri, np.arange(N) - nE - nTri)
mate = newid[mate_full[sid]]
dA = As.diagonal()
has = mate >= 0
off = np.zeros(ns, complex)
Asc = As.tocsr()
idx = np.nonzero(has)[0]
off[idx] = np.asarray(Asc[idx, mate[idx]]).ravel()
def blockjac(r):
    z = r / dA
    i = idx; m = mate[idx]
    det = dA[i] * dA[m] - off[i] * off[i]
    z2 = z.copy()
    z2[i] = (dA[m] * r[i] - off[i] * r[m]) / det
    return z2

dL2 = 1.0 / L2.diagonal(); rhoL2 = None
from emerge_b200.amg import _rho_DinvA
rhoL2 = _rho_DinvA(L2, dL2)
dAw = (Ps.T @ As @ Ps).diagonal()
Aw = (Ps.T @ As @ Ps).tocsr()
print("rho L2", rhoL2, flush=True)

def mg_p2(b, nu=1):
    om = 4.0 / (3.0 * rhoL2)
    x = om * dL2 * b
    rc = I21s.T @ (b - L2 @ x)
    x = x + I21s @ vcycle(H1, rc)
    x = x + om * dL2 * (b - L2 @ x)
    return x

lu_cache = {}
def exact(name, M):
    if name not in lu_cache: lu_cache[name] = spla.splu(sp.csc_matrix(M))
    return lu_cache[name].solve

def prec(variant):
    if variant == "jac":       # what the GPU has today (additive Jacobi on every space)
        PG1 = (Ps @ (G1[~badP][:, ~badNode])).tocsr() if False else None
        dG = (Gs.T @ As @ Gs).diagonal()
        return lambda r: blockjac(r) + Gs @ ((Gs.T @ r) / dG) + Ps @ ((Ps.T @ r) / dAw)
    if variant == "exact":     # ideal two-level: exact solves on the gradient and Whitney spaces
        sg = exact("g", Gs.T @ As @ Gs); sw = exact("w", Aw)
        return lambda r: blockjac(r) + Gs @ sg(Gs.T @ r) + Ps @ sw(Ps.T @ r)
    if variant == "mgexactw":  # MG on gradients, exact Whitney
        sw = exact("w", Aw)
        return lambda r: blockjac(r) + Gs @ (-mg_p2(Gs.T @ r) / k0 ** 2) + Ps @ sw(Ps.T @ r)
    if variant == "mg":        # the design: MG on P2 gradients; Whitney: Jacobi + nodal AMG
        def ap(r):
            z = blockjac(r) + Gs @ (-mg_p2(Gs.T @ r) / k0 ** 2)
            rw = Ps.T @ r
            xw = rw / dAw + Pis @ vcycle(Hn, Pis.T @ rw)
            return z + Ps @ xw
        return ap
    if variant == "mg2":       # same, the Whitney level multiplicative: smooth, nodal correction, smooth
        def ap(r):
            z = blockjac(r) + Gs @ (-mg_p2(Gs.T @ r) / k0 ** 2)
            rw = Ps.T @ r
            xw = 0.7 * rw / dAw
            xw = xw + Pis @ vcycle(Hn, Pis.T @ (rw - Aw @ xw))
            xw = xw + 0.7 * (rw - Aw @ xw) / dAw
            return z + Ps @ xw
        return ap
    if variant.startswith("cyc"):   # cycNM: N gradient cycles, M Whitney cycles (multiplicative Richardson), nodal = scalar Laplace
        ng, nw = int(variant[3]), int(variant[4])
        G1s = G1[~badP][:, ~badNode].tocsr()
        Ln = (G1s.T @ Mw @ G1s).tocsr()
        Hs = sa_hierarchy(Ln)
        Pc = [Pis[:, c::3].tocsr() for c in range(3)]
        AwS = ((Aw + Aw.T) * 0.5).tocsr()
        def nodal(rw):
            return sum(Pc[c] @ vcycle(Hs, Pc[c].T @ rw) for c in range(3))
        def wcycle(rw):
            xw = 0.7 * rw / dAw
            rr = rw - AwS @ xw
            xw = xw + nodal(rr) + G1s @ (-vcycle(Hs, G1s.T @ rr) / k0 ** 2)
            xw = xw + 0.7 * (rw - AwS @ xw) / dAw
            return xw
        def ap(r):
            rg = Gs.T @ r
            xg = mg_p2(rg)
            for _ in range(ng - 1): xg = xg + mg_p2(rg - L2 @ xg)
            z = blockjac(r) + Gs @ (-xg / k0 ** 2)
            rw = Ps.T @ r
            xw = wcycle(rw)
            for _ in range(nw - 1): xw = xw + wcycle(rw - AwS @ xw)
            return z + Ps @ xw
        return ap
    if variant.startswith("helm"):   # nodal branch on the scalar Helmholtz operators Ln - k0^2 Mn (indefinite coarse solve)
        G1s = G1[~badP][:, ~badNode].tocsr()
        Ln = (G1s.T @ Mw @ G1s).tocsr()
        # P1 mass matrix on kept nodes
        tv = np.asarray(t.tets); X = t.nodes[:, tv]                  # (3,4,nT)
        e1, e2, e3 = X[:, 1] - X[:, 0], X[:, 2] - X[:, 0], X[:, 3] - X[:, 0]
        vol = np.abs(np.einsum('it,it->t', e1, np.cross(e2.T, e3.T).T)) / 6
        ii = np.repeat(tv, 4, axis=0); jj = np.tile(tv, (4, 1))
        vv = np.tile(vol / 20, (16, 1)) * (1 + (ii == jj))
        Mn = sp.coo_matrix((vv.ravel(), (ii.ravel(), jj.ravel())), shape=(nN, nN)).tocsr()[~badNode][:, ~badNode].tocsr()
        cs = int(variant[4:]) if len(variant) > 4 else 600
        Hs = sa_hierarchy(Ln, coarse_size=cs)
        Ml = Mn
        for lev in Hs:
            lev['H'] = (lev['A'] - k0 ** 2 * Ml).tocsr()
            lev['hd'] = 1.0 / lev['H'].diagonal()
            if lev['P'] is not None: Ml = (lev['P'].T @ Ml @ lev['P']).tocsr()
        print("helm levels", [l['A'].shape[0] for l in Hs], "min diag sign", [float(np.min(l['H'].diagonal())) > 0 for l in Hs], flush=True)
        def vh(b, lvl=0):
            lev = Hs[lvl]; H = lev['H']
            if lev['P'] is None:
                if 'hlu' not in lev: lev['hlu'] = spla.splu(sp.csc_matrix(H))
                return lev['hlu'].solve(b.real) + 1j * lev['hlu'].solve(b.imag)
            om = 4.0 / (3.0 * lev['rho'])
            x = om * lev['hd'] * b
            x = x + lev['P'] @ vh(lev['P'].T @ (b - H @ x), lvl + 1)
            return x + om * lev['hd'] * (b - H @ x)
        Pc = [Pis[:, c::3].tocsr() for c in range(3)]
        AwS = ((Aw + Aw.T) * 0.5).tocsr()
        def ap(r):
            z = blockjac(r) + Gs @ (-mg_p2(Gs.T @ r) / k0 ** 2)
            rw = Ps.T @ r
            xw = 0.7 * rw / dAw
            rr = rw - AwS @ xw
            xw = xw + sum(Pc[c] @ vh(Pc[c].T @ rr) for c in range(3)) + G1s @ (-vcycle(Hs, G1s.T @ rr) / k0 ** 2)
            xw = xw + 0.7 * (rw - AwS @ xw) / dAw
            return z + Ps @ xw
        return ap
    if variant.startswith("dg"):   # diagnostics: exact solves substituted for individual pieces of the additive variant
        G1s = G1[~badP][:, ~badNode].tocsr()
        Ln = (G1s.T @ Mw @ G1s).tocsr()
        Hs = sa_hierarchy(Ln)
        dG = (Gs.T @ As @ Gs).diagonal()
        R3 = (Ps @ G1s).tocsr()
        Rc = [(Ps @ Pis[:, c::3]).tocsr() for c in range(3)]
        Rn = sp.hstack(Rc).tocsr()
        which = variant[2:]
        sg = exact("g", Gs.T @ As @ Gs) if "g" in which else None
        s3 = exact("r3", R3.T @ As @ R3) if "p" in which else None
        sn = exact("rn", Rn.T @ As @ Rn) if "n" in which else None
        if "l" in which:
            _lu = spla.splu(sp.csc_matrix(Ln))
            sl = lambda b: _lu.solve(np.ascontiguousarray(b.real)) + 1j * _lu.solve(np.ascontiguousarray(b.imag))
        else:
            sl = None
        def ap(r):
            z = blockjac(r) + Ps @ ((Ps.T @ r) / dAw)
            z = z + (Gs @ sg(Gs.T @ r) if sg else Gs @ ((Gs.T @ r) / dG))
            if s3: z = z + R3 @ s3(R3.T @ r)
            elif sl: z = z + R3 @ (-sl(R3.T @ r) / k0 ** 2)
            else: z = z + R3 @ (-vcycle(Hs, R3.T @ r) / k0 ** 2)
            if sn: z = z + Rn @ sn(Rn.T @ r)
            elif sl:
                for c in range(3): z = z + Rc[c] @ sl(Rc[c].T @ r)
            else:
                for c in range(3): z = z + Rc[c] @ vcycle(Hs, Rc[c].T @ r)
            return z
        return ap
    if variant.startswith("padd"):   # additive, port surface term included in the P1-gradient operator (real surrogate |gamma|)
        G1s = G1[~badP][:, ~badNode].tocsr()
        Ln = (G1s.T @ Mw @ G1s).tocsr()
        Ssum = sum(p['S'] for p in S['ports']).tocsr()[sid][:, sid].real.tocsr()
        R3 = (Ps @ G1s).tocsr()
        T = (R3.T @ Ssum @ R3).tocsr()
        beta = np.sqrt(k0 ** 2 - (np.pi / S['a']) ** 2)
        Hg = sa_hierarchy((Ln + (beta / k0 ** 2) * T).tocsr())
        Hs = sa_hierarchy(Ln)
        print("T nnz", T.nnz, "levels", [l['A'].shape[0] for l in Hg], flush=True)
        dG = (Gs.T @ As @ Gs).diagonal()
        Rc = [(Ps @ Pis[:, c::3]).tocsr() for c in range(3)]
        def ap(r):
            z = blockjac(r) + Gs @ ((Gs.T @ r) / dG) + Ps @ ((Ps.T @ r) / dAw)
            z = z + R3 @ (-vcycle(Hg, R3.T @ r) / k0 ** 2)
            for c in range(3): z = z + Rc[c] @ vcycle(Hs, Rc[c].T @ r)
            return z
        return ap
    if variant.startswith("add"):   # purely additive: Jacobi on G and P spaces, AMG V-cycles on P1 gradients and nodal components
        G1s = G1[~badP][:, ~badNode].tocsr()
        Ln = (G1s.T @ Mw @ G1s).tocsr()
        Hs = sa_hierarchy(Ln)
        dG = (Gs.T @ As @ Gs).diagonal()
        R3 = (Ps @ G1s).tocsr()
        Rc = [(Ps @ Pis[:, c::3]).tocsr() for c in range(3)]
        nu = 2 if variant.endswith("2") else 1
        def ap(r):
            z = blockjac(r) + Gs @ ((Gs.T @ r) / dG) + Ps @ ((Ps.T @ r) / dAw)
            z = z + R3 @ (-vcycle(Hs, R3.T @ r, nu=nu) / k0 ** 2)
            for c in range(3): z = z + Rc[c] @ vcycle(Hs, Rc[c].T @ r, nu=nu)
            return z
        return ap
    if variant in ("hx", "hxg", "hx2"):   # Whitney branch: Jacobi + Pi * (scalar Laplace V-cycle per component) * Pi^T
        G1s = G1[~badP][:, ~badNode].tocsr()
        Ln = (G1s.T @ Mw @ G1s).tocsr()
        Hs = sa_hierarchy(Ln)
        print("Ln levels", [l['A'].shape[0] for l in Hs], flush=True)
        Pc = [Pis[:, c::3].tocsr() for c in range(3)]
        def nodal(rw):
            return sum(Pc[c] @ vcycle(Hs, Pc[c].T @ rw) for c in range(3))
        def ap(r):
            z = blockjac(r) + Gs @ (-mg_p2(Gs.T @ r) / k0 ** 2)
            rw = Ps.T @ r
            if variant == "hx2":
                xw = 0.7 * rw / dAw
                xw = xw + nodal(rw - Aw @ xw)
                xw = xw + 0.7 * (rw - Aw @ xw) / dAw
            else:
                xw = rw / dAw + nodal(rw)
                if variant == "hxg": xw = xw + G1s @ (-vcycle(Hs, G1s.T @ rw) / k0 ** 2)
            return z + Ps @ xw
        return ap
    raise ValueError(variant)

def cocr(Aop, b, Minv, rtol=1e-8, maxit=6000):
    x = np.zeros_like(b); r = b.copy(); z = Minv(r); p = z.copy(); Az = Aop @ z; Ap = Az.copy()
    zAz = z @ Az; bn = np.linalg.norm(b); hist = []
    for it in range(1, maxit + 1):
        MAp = Minv(Ap)
        alpha = zAz / (Ap @ MAp)
        x += alpha * p; r -= alpha * Ap; z -= alpha * MAp
        rn = np.linalg.norm(r) / bn; hist.append(rn)
        if rn <= rtol: break
        Az = Aop @ z
        znew = z @ Az
        beta = znew / zAz; zAz = znew
        p = z + beta * p; Ap = Az + beta * Ap
    return x, it, hist

def pcg(Aop, b, Minv, rtol=1e-8, maxit=500):
    x = np.zeros_like(b); r = b.copy(); z = Minv(r); p = z.copy(); rz = np.vdot(r, z); bn = np.linalg.norm(b)
    for it in range(1, maxit + 1):
        Ap = Aop @ p; a = rz / np.vdot(p, Ap); x += a * p; r -= a * Ap
        if np.linalg.norm(r) / bn <= rtol: return it
        z = Minv(r); rzn = np.vdot(r, z); p = z + (rzn / rz) * p; rz = rzn
    return maxit

if "checkamg" in variants:
    variants.remove("checkamg")
    rng = np.random.default_rng(1)
    for name, M, pre in (("L1 vcycle", L1, lambda r: vcycle(H1, r)), ("L2 two-grid", L2, mg_p2),
                         ("An vcycle", An, lambda r: vcycle(Hn, r)),
                         ("Kw+tauMw jac+nodal", (Kw + tau * Mw).tocsr(), None)):
        b = M @ (rng.standard_normal(M.shape[0]) + 0j)
        if pre is None:
            dd = M.diagonal()
            G1s = G1[~badP][:, ~badNode].tocsr()
            pre = lambda r: r / dd + Pis @ vcycle(Hn, Pis.T @ r) + G1s @ (vcycle(H1x, G1s.T @ r) / tau)
            H1x = sa_hierarchy((G1s.T @ Mw @ G1s).tocsr())
        print(f"PCG {name:22s} n={M.shape[0]:8d} iterations {pcg(M, b, pre)}", flush=True)

for v in variants:
    t0 = time.time()
    Minv = prec(v)
    x, it, hist = cocr(A if "--trueA" in sys.argv else As, rhs[0], Minv, rtol=float(sys.argv[sys.argv.index("--rtol") + 1]) if "--rtol" in sys.argv else 1e-8)
    print("true residual on A:", np.linalg.norm(rhs[0] - A @ x) / np.linalg.norm(rhs[0]))
    print(f"variant {v:10s} iterations {it:5d}  final {hist[-1]:.2e}  true {np.linalg.norm(rhs[0]-As@x)/np.linalg.norm(rhs[0]):.2e} ({time.time()-t0:.1f}s)  hist@[10,50,100,200]: "
          + " ".join(f"{hist[min(k, len(hist)-1)]:.1e}" for k in (10, 50, 100, 200)), flush=True)
