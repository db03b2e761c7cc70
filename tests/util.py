"""Shared helpers of the test-suite: golden fixtures -> mesh tables / boundary conditions."""
import os

import numpy as np
import scipy.sparse as sp

from emerge_b200 import bc as B
from emerge_b200.synthmesh import mesh_tables, tri_ids_of

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
    t = mesh_tables(np.ascontiguousarray(g["nodes"].T), np.ascontiguousarray(g["tets"].T.astype(np.int64)),
                    g["edges"].astype(np.int64), g["tris"].astype(np.int64))
    return g, t


def tag_tris(g, t, tag):
    return tri_ids_of(t, g["face_tris"][g["face_tag"] == tag].astype(np.int64))


def golden_bcs(g, t):
    kind = str(g["kind"])
    if kind == "rectwg":
        a, b, L = g["dims"]
        pec = B.PEC(np.concatenate([tag_tris(g, t, k) for k in (1, 2, 3, 4)]))
        p1 = B.RectangularWaveguide(tag_tris(g, t, 5), 1, B.CoordSys(origin=(0, 0, 0.0)), (a, b))
        p2 = B.RectangularWaveguide(tag_tris(g, t, 6), 2, B.CoordSys(origin=(0, 0, L)), (a, b))
        return [pec, p1, p2]
    if kind == "abc_lumped":
        pec = B.PEC(np.concatenate([tag_tris(g, t, k) for k in (5, 7)]))
        w, h, z0 = g["lumped"]
        basis = g["port_cs_basis"]
        cs = B.CoordSys(basis[:, 0], basis[:, 1], basis[:, 2], g["port_cs_origin"])
        lp = B.LumpedPort(tag_tris(g, t, 8), 1, cs, w, h, (0, 0, 1), Z0=z0)
        abc = B.AbsorbingBoundary(np.concatenate([tag_tris(g, t, k) for k in (1, 2, 3, 4, 6)]))
        return [pec, lp, abc]
    if kind == "modal_microstrip":
        a, b, L = g["dims"]
        pec = B.PEC(np.concatenate([tag_tris(g, t, k) for k in (1, 2, 3, 4, 7)]))
        ports = []
        for n, (tag, z) in enumerate([(5, 0.0), (6, L)], start=1):
            beta, k0m, tem, fm = g[f"mode_scalars_p{n}"][:4]
            ports.append(B.ModalPort(tag_tris(g, t, tag), n, B.CoordSys(origin=(0, 0, z)),
                                     B.SampledField(g[f"mode_pts_p{n}"], g[f"mode_E_p{n}"]), beta, k0m, TEM=bool(tem),
                                     freq_mode=fm, modetype=str(g[f"mode_type_p{n}"])))
        return [pec] + ports
    raise ValueError(kind)


def csr(g, prefix, n, data_key=None):
    return sp.csr_matrix((g[data_key or prefix + "_data"], g[prefix + "_indices"], g[prefix + "_indptr"]), shape=(n, n))


def db_deg_close(S, Sref, db_tol=1e-3, deg_tol=0.1, floor=1e-4):
    """|S| within db_tol dB and angle within deg_tol degrees; entries below `floor` compared absolutely."""
    S, Sref = np.asarray(S).ravel(), np.asarray(Sref).ravel()
    ok = True
    for a, b in zip(S, Sref):
        if abs(b) < floor:
            ok &= abs(a - b) < floor * 1e-2 + 1e-9
            continue
        ok &= abs(20 * np.log10(abs(a) / abs(b))) < db_tol
        ok &= abs(np.angle(a / b, deg=True)) < deg_tol
    return bool(ok)


def oracle_system(g, t, k0, E, Bm):
    """K(f) = E - k0^2 B + sum gamma_p S_p and the port right-hand sides from the numpy oracle (TEST INFRASTRUCTURE),
    following Assembler.assemble_freq_matrix (fem/physics/edm/assembler.py:312-388)."""
    from oracle import ned2_oracle as O
    from emerge_b200.sweep import dunavant4
    N = t.n_field
    K = E - Bm * k0 ** 2
    DP = dunavant4()
    bvecs = {}
    for bc in golden_bcs(g, t):
        if not hasattr(bc, "get_gamma"):
            continue
        ids = bc.tri_ids
        v = t.nodes[:, t.tris[:, ids]]                                  # (3, 3 verts, n)
        if bc._include_force:
            loc = np.einsum("ij,jvn->ivn", bc.get_inv_basis(), v - bc.cs.origin[:, None, None])
            x, y = loc[0].T, loc[1].T
            S = O.tri_surface_matrix(x, y)
            xq, yq = x @ DP[1:4], y @ DP[1:4]
            U = bc.get_Uinc(xq.T.ravel(), yq.T.ravel(), k0).reshape(3, 6, len(ids))
            bl = O.tri_forcing(x, y, U[:2])
            bv = np.zeros(N, dtype=complex)
            np.add.at(bv, t.tri_to_field[:, ids].T, bl)
            bvecs[bc.port_number] = bv
        else:
            x, y = O.abc_local_frame(v.transpose(2, 1, 0))
            S = O.tri_surface_matrix(x, y, t.edge_lengths[t.tri_to_edge[:, ids]].T)
        K = K + O.gen_csr_tri(N, t.tri_to_field, ids, bc.get_gamma(k0) * S)
    return K, bvecs
