"""Prototype (CPU, scratch): p-multigrid style preconditioner - the order-1 (Whitney) problem solved by an inner Krylov
iteration, outer flexible GMRES on the true A.  usage: proto_pmg.py nx ny nz [inner iterations ...]  (-1 = exact)"""
import sys, time
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
from proto_common import *
from emerge_b200.auxspace import build_aux_spaces, nodal_interpolation
from emerge_b200.amg import sa_hierarchy, _rho_DinvA

nx, ny, nz = [int(v) for v in sys.argv[1:4]]
inners = [int(v) for v in sys.argv[4:]] or [0, 5, 10, 20]
f = 10e9
L = nz * 22.86e-3 / nx
S = waveguide_system(nx, ny, nz, L)
t = S['t']
A, rhs = system_at(S, f)
A = A.tocsr()
As = ((A + A.T) * 0.5).tocsr()
k0 = 2 * np.pi * f / 299792458
sid = S['solve_ids']; N = S['N']; ns = len(sid)
print(f"tets {t.tets.shape[1]} ns {ns} nnz {A.nnz}", flush=True)
Ms = S['B'].tocsr()[sid][:, sid].real.tocsr(); Ms = ((Ms + Ms.T) * 0.5).tocsr()
G, P, G1 = build_aux_spaces(t)
nN, nE, nTri = t.nodes.shape[1], t.edges.shape[1], t.tris.shape[1]
keep = np.zeros(N, bool); keep[sid] = True; elim = ~keep
def restrict(R):
    bad = np.asarray(abs(R[elim]).sum(axis=0)).ravel() > 0
    return R[keep][:, ~bad].tocsr(), bad
Gs, badG = restrict(G)
Ps, badP = restrict(P)
badN = np.asarray(abs(G1[badP]).sum(axis=0)).ravel() > 0
G1s = G1[~badP][:, ~badN].tocsr()
Pc = [p[~badP][:, ~badN].tocsr() for p in nodal_interpolation(t)]
Aw = (Ps.T @ As @ Ps).tocsr()
Mw = (Ps.T @ Ms @ Ps).tocsr()
dAw = Aw.diagonal()
Ln = (G1s.T @ Mw @ G1s).tocsr()
Hs = sa_hierarchy(Ln)
# gradient space: P2 Laplacian, two-grid with P1 AMG
L2 = (Gs.T @ Ms @ Gs).tocsr()
e = np.asarray(t.edges)
I21 = sp.coo_matrix((np.concatenate([np.ones(nN), 0.5 * np.ones(nE), 0.5 * np.ones(nE)]),
                     (np.concatenate([np.arange(nN), nN + np.arange(nE), nN + np.arange(nE)]),
                      np.concatenate([np.arange(nN), e[0], e[1]]))), shape=(nN + nE, nN)).tocsr()
badN2 = np.asarray(abs(I21[badG]).sum(axis=0)).ravel() > 0
I21s = I21[~badG][:, ~badN2].tocsr()
H1 = sa_hierarchy((I21s.T @ L2 @ I21s).tocsr())
dL2 = 1.0 / L2.diagonal(); rhoL2 = _rho_DinvA(L2, dL2)
dG = (Gs.T @ As @ Gs).diagonal()
print(f"Whitney {Aw.shape[0]} nnz/row {Aw.nnz / Aw.shape[0]:.1f}; nodes {Ln.shape[0]}; P2 {L2.shape[0]}", flush=True)

def vcycle(H, b, lvl=0):
    lev = H[lvl]; Al = lev['A']
    if lev['P'] is None:
        if 'lu' not in lev: lev['lu'] = spla.splu(sp.csc_matrix(Al + 1e-12 * sp.identity(Al.shape[0])))
        return lev['lu'].solve(np.ascontiguousarray(b.real)) + 1j * lev['lu'].solve(np.ascontiguousarray(b.imag))
    om = 4.0 / (3.0 * lev['rho'])
    x = om * lev['dinv'] * b
    x = x + lev['P'] @ vcycle(H, lev['P'].T @ (b - Al @ x), lvl + 1)
    return x + om * lev['dinv'] * (b - Al @ x)
def mg_p2(b):
    om = 4.0 / (3.0 * rhoL2)
    x = om * dL2 * b
    x = x + I21s @ vcycle(H1, I21s.T @ (b - L2 @ x))
    return x + om * dL2 * (b - L2 @ x)

# fine-level block Jacobi
newid = -np.ones(N, int); newid[sid] = np.arange(ns)
mate_full = np.where(np.arange(N) < nE + nTri, np.arange(N) + nE + nTri, np.arange(N) - nE - nTri)
mate = newid[mate_full[sid]]
dA = As.diagonal(); idx = np.nonzero(mate >= 0)[0]
off = np.zeros(ns, complex); off[idx] = np.asarray(As[idx, mate[idx]]).ravel()
def blockjac(r):
    z = r / dA
    i = idx; m = mate[idx]
    det = dA[i] * dA[m] - off[i] * off[i]
    z[i] = (dA[m] * r[i] - off[i] * r[m]) / det
    return z

def whit_prec(rw):       # what the GPU applies on the Whitney space today (additive)
    x = rw / dAw + G1s @ (-vcycle(Hs, G1s.T @ rw) / k0 ** 2)
    for c in range(3): x = x + Pc[c] @ vcycle(Hs, Pc[c].T @ rw)
    return x
stats = dict(inner=0)
def whit_solve(rw, k):
    if k == 0: return whit_prec(rw)
    if k < 0:
        if 'lu' not in stats: stats['lu'] = spla.splu(sp.csc_matrix(Aw))
        return stats['lu'].solve(rw)
    # k COCR iterations on Aw with the additive preconditioner
    x = np.zeros_like(rw); r = rw.copy(); z = whit_prec(r); p = z.copy(); Az = Aw @ z; Ap = Az.copy(); zAz = z @ Az
    for it in range(k):
        MAp = whit_prec(Ap); alpha = zAz / (Ap @ MAp)
        x += alpha * p; r -= alpha * Ap; z -= alpha * MAp
        Az = Aw @ z; zn = z @ Az; beta = zn / zAz; zAz = zn
        p = z + beta * p; Ap = Az + beta * Ap
        stats['inner'] += 1
    return x

GRAD = "jac" if "--gjac" in sys.argv else "mg"
def make_prec(k):
    def ap(r):
        z = blockjac(r)
        if GRAD == "mg": z = z + Gs @ (-mg_p2(Gs.T @ r) / k0 ** 2)
        else: z = z + Gs @ ((Gs.T @ r) / dG)
        return z + Ps @ whit_solve(Ps.T @ r, k)
    return ap

def fgmres(Aop, b, Minv, rtol=1e-8, restart=60, maxit=3000):
    x = np.zeros_like(b); bn = np.linalg.norm(b); its = 0; hist = []
    while its < maxit:
        r = b - Aop @ x; beta = np.linalg.norm(r)
        if beta / bn <= rtol: break
        V = [r / beta]; Z = []; H = np.zeros((restart + 1, restart), complex); g = np.zeros(restart + 1, complex); g[0] = beta
        for j in range(restart):
            z = Minv(V[j]); Z.append(z); w = Aop @ z
            for i in range(j + 1):
                H[i, j] = np.vdot(V[i], w); w = w - H[i, j] * V[i]
            H[j + 1, j] = np.linalg.norm(w); V.append(w / H[j + 1, j]); its += 1
            y, res, *_ = np.linalg.lstsq(H[:j + 2, :j + 1], g[:j + 2], rcond=None)
            rn = np.linalg.norm(g[:j + 2] - H[:j + 2, :j + 1] @ y) / bn; hist.append(rn)
            if rn <= rtol or its >= maxit: break
        x = x + sum(y[i] * Z[i] for i in range(len(y)))
    return x, its, hist

for k in inners:
    stats['inner'] = 0
    t0 = time.time()
    x, it, hist = fgmres(A, rhs[0], make_prec(k))
    tr = np.linalg.norm(rhs[0] - A @ x) / np.linalg.norm(rhs[0])
    print(f"inner {k:3d}: outer FGMRES(60) iterations {it:5d}  inner total {stats['inner']:6d}  true relres {tr:.2e}  ({time.time()-t0:.0f}s)"
          f"  hist@[10,20,40,80]: " + " ".join(f"{hist[min(q, len(hist)-1)]:.1e}" for q in (10, 20, 40, 80)), flush=True)
