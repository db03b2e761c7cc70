"""Auxiliary spaces for the multilevel preconditioner (host side, numpy; one-time per mesh).

The second-order Nedelec space of the reference contains
  * the gradients of the quadratic Lagrange space  (the null-space of the curl-curl matrix E), and
  * the lowest-order (Whitney) Nedelec space.
This module builds the two sparse transfer matrices in the reference's dof numbering,
    G  (n_field x (nN + nE)) : coefficients of grad(phi_k), phi_k the P2 nodal functions
                               (vertex functions lam_v(2 lam_v - 1), edge functions 4 lam_A lam_B),
    P  (n_field x nE)        : coefficients of the Whitney functions w_AB = lam_B grad lam_A - lam_A grad lam_B,
and G1 (nE x nN), the P1 gradient in the Whitney basis.  They are used by the CUDA solver as an additive
multilevel preconditioner  M^-1 = D^-1 + G D_G^-1 G^T + P (D_P^-1 + G1 D_G1^-1 G1^T) P^T.

Derivation of the entries (no quadrature, no per-tet work): the dofs of the reference's basis are point
functionals of the tangential field,
    c_edge-a(A,B) = -E(v_A).t_AB,  c_edge-b(A,B) = -E(v_B).t_AB            (t_AB = (v_B - v_A)/l_AB),
and, on a face (A,B,E) with centroid c, the two face coefficients follow from E_t(c):
    u = c_fa l_AE,  w = c_fb l_AB:   u - 2w = R1,  2u - w = R2,
    R1 = 9 E(c).d_AB + 2 s_AB - s_BE + s_AE,   R2 = 9 E(c).d_AE + s_AB + s_BE + 2 s_AE,   s_PQ = -(E_P + E_Q).d_PQ,
because at the centroid every edge function of edge (P,Q) equals l_PQ (g_P - g_Q)/9, face-a = -l_AE (g_A - g_E)/9,
face-b = l_AB (g_A - g_B)/9, and g_P.d_QR = delta_PR - delta_PQ.  All targets are linear fields, so only their
vertex values dotted with edge vectors enter; these are integers (delta combinations), hence every entry of G and P
is (small integer)/(edge length).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def _gdot(V, Q, R):
    """grad(lam_V) . (v_R - v_Q) = delta_VR - delta_VQ (local indices 0..2)."""
    return float(V == R) - float(V == Q)


def _face_tables():
    """For the 6 P2 targets (3 vertex, 3 edge fns) and 3 Whitney targets on a face with local vertices (0,1,2) =
    (A,B,E) and local edges e0=(0,1), e1=(1,2), e2=(0,2): the pair (u, w) = (c_fa*l_AE, c_fb*l_AB)."""
    edges = [(0, 1), (1, 2), (0, 2)]

    def solve(Edot):           # Edot(u, Q, R) = E_u . (v_R - v_Q)
        s = {e: -(Edot(e[0], *e) + Edot(e[1], *e)) for e in edges}
        Ec_AB = sum(Edot(u, 0, 1) for u in range(3)) / 3.0
        Ec_AE = sum(Edot(u, 0, 2) for u in range(3)) / 3.0
        R1 = 9 * Ec_AB + 2 * s[(0, 1)] - s[(1, 2)] + s[(0, 2)]
        R2 = 9 * Ec_AE + s[(0, 1)] + s[(1, 2)] + 2 * s[(0, 2)]
        return (2 * R2 - R1) / 3.0, (R2 - 2 * R1) / 3.0

    vert = [solve(lambda u, Q, R, V=V: (4.0 * (u == V) - 1.0) * _gdot(V, Q, R)) for V in range(3)]
    edge = [solve(lambda u, Q, R, P_=P_, Q_=Q_: 4.0 * ((u == P_) * _gdot(Q_, Q, R) + (u == Q_) * _gdot(P_, Q, R)))
            for (P_, Q_) in edges]
    whit = [solve(lambda u, Q, R, P_=P_, Q_=Q_: (u == Q_) * _gdot(P_, Q, R) - (u == P_) * _gdot(Q_, Q, R))
            for (P_, Q_) in edges]
    return np.array(vert), np.array(edge), np.array(whit)


def _rows_to_csr(rows, cols, vals, blocks, shape):
    """CSR straight from per-row-block entry lists: every block of rows (edge-a, face-a, edge-b, face-b: contiguous,
    ascending dof ranges that tile [0, N)) has a fixed number of entries per row, so no COO sort is needed.
    rows/cols/vals: parallel lists of arrays, one entry per row of the block the row array belongs to."""
    ip_parts, ix_parts, dv_parts = [], [], []
    offset = 0
    for blk in blocks:
        sel = [k for k, r in enumerate(rows) if r is blk]
        n, m = len(blk), len(sel)
        ix = np.empty((n, m), dtype=np.int32)
        dv = np.empty((n, m), dtype=np.float64)
        for j, k in enumerate(sel):
            ix[:, j] = cols[k]
            dv[:, j] = vals[k]
        ip_parts.append(offset + m * np.arange(n, dtype=np.int64))
        ix_parts.append(ix.ravel())
        dv_parts.append(dv.ravel())
        offset += n * m
    indptr = np.concatenate(ip_parts + [np.array([offset], dtype=np.int64)])
    M = sp.csr_matrix((np.concatenate(dv_parts), np.concatenate(ix_parts), indptr), shape=shape)
    M.eliminate_zeros()
    M.sort_indices()
    return M


def build_aux_spaces(tables):
    """tables: MeshTables-like (nodes (3,nN), edges (2,nE), tris (3,nTri), tri_to_edge (3,nTri), edge_lengths).
    Returns (G, P, G1) as scipy CSR in the reference's dof numbering
    [edge-a | face-a | edge-b | face-b] (fem/elements/nedelec2.py:46-62)."""
    nodes = np.asarray(tables.nodes)
    edges = np.asarray(tables.edges)
    tris = np.asarray(tables.tris)
    t2e = np.asarray(tables.tri_to_edge)
    nN, nE, nTri = nodes.shape[1], edges.shape[1], tris.shape[1]
    N = 2 * nE + 2 * nTri
    d = nodes[:, edges[1]] - nodes[:, edges[0]]
    ell = np.sqrt((d ** 2).sum(axis=0))
    il = 1.0 / ell
    ea, eb = np.arange(nE), np.arange(nE) + nE + nTri
    fa, fb = np.arange(nTri) + nE, np.arange(nTri) + 2 * nE + nTri
    A, Bv = edges[0], edges[1]
    ecol = nN + np.arange(nE)
    # edge rows of G: c_a = -grad(phi)(v_A).t, c_b = -grad(phi)(v_B).t
    rows = [ea, ea, ea, eb, eb, eb]
    cols = [A, Bv, ecol, A, Bv, ecol]
    vals = [3 * il, 1 * il, -4 * il, -1 * il, -3 * il, 4 * il]
    # face rows
    vt, et, wt = _face_tables()
    lAE = ell[t2e[2]]         # tri edges (1-2, 2-3, 1-3) = (A,B),(B,E),(A,E)  (fem/mesh3d.py:330-335)
    lAB = ell[t2e[0]]
    for k in range(3):        # vertex targets
        rows += [fa, fb]
        cols += [tris[k], tris[k]]
        vals += [vt[k, 0] / lAE, vt[k, 1] / lAB]
    for k in range(3):        # edge targets
        rows += [fa, fb]
        cols += [nN + t2e[k], nN + t2e[k]]
        vals += [et[k, 0] / lAE, et[k, 1] / lAB]
    G = _rows_to_csr(rows, cols, vals, (ea, fa, eb, fb), (N, nN + nE))
    # Whitney prolongation
    rows = [ea, eb]
    cols = [ea, ea]
    vals = [il, il]
    for k in range(3):
        rows += [fa, fb]
        cols += [t2e[k], t2e[k]]
        vals += [wt[k, 0] / lAE, wt[k, 1] / lAB]
    P = _rows_to_csr(rows, cols, vals, (ea, fa, eb, fb), (N, nE))
    # P1 gradient in the Whitney basis: grad(lam_v) = sum_e G1[e,v] w_e ; w_AB.t_AB = -1/l  =>  G1[e,A]=+1, G1[e,B]=-1
    G1 = sp.coo_matrix((np.concatenate([np.ones(nE), -np.ones(nE)]),
                        (np.concatenate([ea, ea]), np.concatenate([A, Bv]))), shape=(nE, nN)).tocsr()
    return G, P, G1


def build_aux_spaces_paired(tables, keep):
    """G and P restricted to the kept dofs with rows already in the library's PAIR order of the solve space (row 2j / 2j+1 =
    first / second function of kept entity j: kept edges in ascending order, then kept faces; csrc/operators.cu) and
    columns touching an eliminated dof dropped - built directly, without the full matrices, row gathers or permutations
    of build_aux_spaces + restriction (same entries; tests/test_auxspace_cpu.py).
    keep: bool (n_field,), identical for the two functions of every entity.  Returns (Gs, Ps, badP, G1)."""
    nodes = np.asarray(tables.nodes)
    edges = np.asarray(tables.edges)
    tris = np.asarray(tables.tris)
    t2e = np.asarray(tables.tri_to_edge)
    nN, nE, nTri = nodes.shape[1], edges.shape[1], tris.shape[1]
    d = nodes[:, edges[1]] - nodes[:, edges[0]]
    ell = np.sqrt((d ** 2).sum(axis=0))
    il = 1.0 / ell
    ke = np.asarray(keep[:nE], dtype=bool)
    kf = np.asarray(keep[nE:nE + nTri], dtype=bool)
    vt, et, wt = _face_tables()
    lAE, lAB = ell[t2e[2]], ell[t2e[0]]

    def edge_rows(sel, which):
        """(cols, vals) of shape (n, 2, m): functions a and b of the selected edges"""
        e = np.nonzero(sel)[0]
        A, B, i = edges[0, e], edges[1, e], il[e]
        if which == "G":
            cols = np.stack([A, B, nN + e], axis=1)
            va = np.stack([3 * i, i, -4 * i], axis=1)
            vb = np.stack([-i, -3 * i, 4 * i], axis=1)
        else:
            cols = e[:, None]
            va = vb = i[:, None]
        return np.stack([cols, cols], axis=1), np.stack([va, vb], axis=1)

    def face_rows(sel, which):
        f = np.nonzero(sel)[0]
        ia, ib = 1.0 / lAE[f], 1.0 / lAB[f]
        if which == "G":
            cols = np.concatenate([tris[:, f].T, nN + t2e[:, f].T], axis=1)                       # (n, 6)
            ta = np.concatenate([vt[:, 0], et[:, 0]])
            tb = np.concatenate([vt[:, 1], et[:, 1]])
        else:
            cols = t2e[:, f].T                                                                    # (n, 3)
            ta, tb = wt[:, 0], wt[:, 1]
        va = ta[None, :] * ia[:, None]
        vb = tb[None, :] * ib[:, None]
        return np.stack([cols, cols], axis=1), np.stack([va, vb], axis=1)

    def assemble(which, ncol):
        bad = np.zeros(ncol, dtype=bool)
        for rows in (edge_rows(~ke, which), face_rows(~kf, which)):      # columns touched by eliminated rows
            c, v = rows
            bad[c[v != 0]] = True
        newcol = np.cumsum(~bad) - 1
        ip_parts, ix_parts, dv_parts, base = [], [], [], 0
        for c, v in (edge_rows(ke, which), face_rows(kf, which)):
            n, _, m = c.shape
            c2, v2 = c.reshape(2 * n, m), v.reshape(2 * n, m)
            ok = (v2 != 0) & ~bad[c2]
            cnt = ok.sum(axis=1)
            ip_parts.append(base + np.concatenate([[0], np.cumsum(cnt)[:-1]]) if 2 * n else np.zeros(0, dtype=np.int64))
            base += int(cnt.sum())
            ix_parts.append(newcol[c2[ok]].astype(np.int32))
            dv_parts.append(v2[ok])
        indptr = np.concatenate(ip_parts + [np.array([base])]).astype(np.int64)
        M = sp.csr_matrix((np.concatenate(dv_parts), np.concatenate(ix_parts), indptr),
                          shape=(len(indptr) - 1, int((~bad).sum())))
        M.sort_indices()
        return M, bad

    Gs, _ = assemble("G", nN + nE)
    Ps, badP = assemble("P", nE)
    ea = np.arange(nE)
    G1 = sp.coo_matrix((np.concatenate([np.ones(nE), -np.ones(nE)]),
                        (np.concatenate([ea, ea]), np.concatenate([edges[0], edges[1]]))), shape=(nE, nN)).tocsr()
    return Gs, Ps, badP, G1


def _edge_rows_csr(edges, nN, va, vb):
    """CSR (nE x nN) with the two entries (A: va, B: vb) per edge row; edges are ascending within a column (A < B), so the
    rows are built sorted, without a COO conversion"""
    nE = edges.shape[1]
    ix = np.empty(2 * nE, dtype=np.int32)
    ix[0::2], ix[1::2] = edges[0], edges[1]
    dv = np.empty(2 * nE, dtype=np.float64)
    dv[0::2], dv[1::2] = va, vb
    M = sp.csr_matrix((dv, ix, 2 * np.arange(nE + 1, dtype=np.int64)), shape=(nE, nN))
    if nE and not np.all(edges[0] < edges[1]):
        M.sort_indices()
    return M


def p1_gradient(tables):
    """G1 (nE x nN): the P1 gradient in the Whitney basis, grad(lam_v) = sum_e G1[e, v] w_e with G1[e, A] = +1, G1[e, B] = -1"""
    edges = np.asarray(tables.edges)
    nN, nE = np.asarray(tables.nodes).shape[1], edges.shape[1]
    return _edge_rows_csr(edges, nN, np.ones(nE), -np.ones(nE))


def nodal_interpolation(tables):
    """Pi_c (nE x nN), c = x, y, z: Whitney coefficients of the nodal vector field e_c * lam_v.  With the reference's sign
    convention w_AB.t_AB = -1/l on edge (A,B), the coefficient of a field E is -(E(v_A) + E(v_B))/2 . (v_B - v_A)."""
    nodes = np.asarray(tables.nodes)
    edges = np.asarray(tables.edges)
    nN = nodes.shape[1]
    d = nodes[:, edges[1]] - nodes[:, edges[0]]
    return [_edge_rows_csr(edges, nN, -0.5 * d[c], -0.5 * d[c]) for c in range(3)]


def p1_stiffness_mass(tables, weight=None):
    """P1 Lagrange stiffness (grad lam_i, w grad lam_j) with a per-tet scalar weight, and the lumped mass vector."""
    nodes = np.asarray(tables.nodes)
    tets = np.asarray(tables.tets)
    nN, nT = nodes.shape[1], tets.shape[1]
    X = nodes[:, tets]                                             # (3, 4, nT)
    e1, e2, e3 = X[:, 1] - X[:, 0], X[:, 2] - X[:, 0], X[:, 3] - X[:, 0]

    def cross(a, b):                                               # rows are contiguous: 6 fused passes instead of np.cross on views
        return np.stack([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])
    g1, g2, g3 = cross(e2, e3), cross(e3, e1), cross(e1, e2)
    det = np.einsum("it,it->t", e1, g1)
    vol = np.abs(det) / 6.0
    w = vol if weight is None else vol * np.asarray(weight, dtype=float)
    t2e = getattr(tables, "tet_to_edge", None)
    if t2e is not None:
        # one value per mesh edge (sum over its tetrahedra) and per node: two bincounts and a duplicate-free COO of
        # 2 nE + nN entries instead of sorting / merging the 16 nT element entries.  grad lam_i = g_i / det: the products
        # are formed from the unscaled g_i and scaled once by w / det^2.
        t2e = np.asarray(t2e)
        edges = np.asarray(tables.edges)
        nE = edges.shape[1]
        g = (-(g1 + g2 + g3), g1, g2, g3)
        s = w / (det * det)
        le = ((0, 1), (0, 2), (0, 3), (1, 2), (3, 1), (2, 3))      # local edge order of tet_to_edge (fem/mesh3d.py:292)
        off = np.stack([np.einsum("xt,xt->t", g[a], g[b]) * s for a, b in le])                 # (6, nT)
        eval_ = np.bincount(t2e.ravel(), weights=off.ravel(), minlength=nE)
        dd = np.stack([np.einsum("xt,xt->t", g[a], g[a]) * s for a in range(4)])               # (4, nT)
        dia = np.bincount(tets.ravel(), weights=dd.ravel(), minlength=nN)
        nd = np.arange(nN)
        L = sp.coo_matrix((np.concatenate([eval_, eval_, dia]),
                           (np.concatenate([edges[0], edges[1], nd]), np.concatenate([edges[1], edges[0], nd]))),
                          shape=(nN, nN)).tocsr()
        mass = np.bincount(tets.ravel(), weights=np.repeat(vol[None, :] / 4.0, 4, axis=0).ravel(), minlength=nN)
        return L, mass
    grad = np.stack([-(g1 + g2 + g3), g1, g2, g3], axis=0) / det   # (4, 3, nT)
    loc = np.einsum("ixt,jxt->ijt", grad, grad) * w                # (4, 4, nT)
    ii = np.broadcast_to(tets[:, None, :], (4, 4, nT))
    jj = np.broadcast_to(tets[None, :, :], (4, 4, nT))
    L = sp.coo_matrix((loc.ravel(), (ii.ravel(), jj.ravel())), shape=(nN, nN)).tocsr()
    mass = np.zeros(nN)
    np.add.at(mass, tets.ravel(), np.repeat(vol[None, :] / 4.0, 4, axis=0).ravel())
    return L, mass
