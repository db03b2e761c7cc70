"""Host logic: the direct construction of the restricted, pair-ordered auxiliary transfer matrices equals the generic
construction followed by restriction and row permutation (emerge_b200/auxspace.py)."""
import numpy as np

from emerge_b200 import auxspace
from emerge_b200.synthmesh import box_mesh, mesh_tables, tri_ids_of


def test_paired_construction_matches_generic_path():
    a, b, L = 22.86e-3, 10.16e-3, 30e-3
    box = box_mesh(5, 3, 7, a, b, L, jitter=0.1, seed=1)
    t = mesh_tables(box.nodes_xyz, box.tets)
    nE, nTri = t.edges.shape[1], t.tris.shape[1]
    N = 2 * nE + 2 * nTri
    tag = lambda k: tri_ids_of(t, box.face_tris[box.face_tag == k])
    pec = np.unique(np.asarray(t.tri_to_field)[:, np.concatenate([tag(k) for k in (1, 2, 3, 4)])].ravel())
    keep = np.ones(N, dtype=bool)
    keep[pec] = False
    H = nE + nTri
    assert np.array_equal(keep[:H], keep[H:])                     # PEC removes both functions of an entity
    # the library's pair order: solve index = 2 * rank inside its half + half (csrc/operators.cu::k_newid)
    kept = np.nonzero(keep)[0]
    rank = np.cumsum(keep) - 1
    internal = np.where(kept < H, 2 * rank[kept], 2 * (rank[kept] - len(kept) // 2) + 1)
    rows_int = np.empty(len(kept), dtype=np.int64)
    rows_int[internal] = kept
    G, P, G1 = auxspace.build_aux_spaces(t)
    elim = ~keep

    def restrict(R):
        bad = np.asarray(abs(R[elim]).sum(axis=0)).ravel() > 0
        return R[rows_int][:, ~bad].tocsr(), bad
    Gref, _ = restrict(G)
    Pref, badPref = restrict(P)
    Gs, Ps, badP, G1b = auxspace.build_aux_spaces_paired(t, keep)
    assert np.array_equal(badP, badPref)
    for A, B in ((Gs, Gref), (Ps, Pref), (G1b, G1)):
        assert A.shape == B.shape
        A.sort_indices(); B.sort_indices()
        B.eliminate_zeros()
        assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
        assert np.array_equal(A.data, B.data)


def test_p1_stiffness_edge_path_matches_element_path():
    """The edge-wise P1 stiffness (two bincounts over the mesh edges, products of the unscaled gradients scaled once) equals
    the element-wise COO assembly, with a per-tet weight, on a jittered mesh; the lumped mass sums to the volume."""
    from types import SimpleNamespace
    a, b, L = 22.86e-3, 10.16e-3, 30e-3
    box = box_mesh(5, 3, 7, a, b, L, jitter=0.15, seed=3)
    t = mesh_tables(box.nodes_xyz, box.tets)
    w = np.random.default_rng(0).uniform(0.5, 2.0, t.tets.shape[1])
    L1, m1 = auxspace.p1_stiffness_mass(t, w)
    L2, m2 = auxspace.p1_stiffness_mass(SimpleNamespace(nodes=t.nodes, tets=t.tets), w)      # no tet_to_edge: element path
    assert abs(L1 - L2).max() <= 1e-13 * abs(L2).max()
    assert np.allclose(m1, m2, rtol=1e-14, atol=0)
    assert abs(m1.sum() - a * b * L) <= 1e-12 * a * b * L
    assert abs(np.asarray(L1.sum(axis=1)).ravel()).max() <= 1e-10 * abs(L1).max()           # constants are in the kernel
