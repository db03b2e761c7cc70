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
