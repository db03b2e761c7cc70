"""CPU: the numpy oracle against the golden vectors generated from the unmodified reference."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import ned2_oracle as O
from tests.util import load_golden, golden_bcs, csr, tag_tris, oracle_system
from emerge_b200.sweep import dunavant4

RTOL = 1e-12      # relative to the largest entry (north_star: K/M to 1e-12 in FP64; reference is fastmath)


def test_element_matrices_vs_reference():
    g, t = load_golden("wg_tiny")
    Em, Bm = O.element_matrices(t.nodes, t.tets, t.edges, t.tris, t.edge_lengths, t.tet_to_field, t.tet_to_edge,
                                g["ur"], g["er"], np.arange(6))
    assert np.abs(Em - g["elemE"]).max() <= RTOL * np.abs(g["elemE"]).max()
    assert np.abs(Bm - g["elemB"]).max() <= RTOL * np.abs(g["elemB"]).max()
    # quirks: K symmetric, M face-face blocks not (fem/mth/tet.py:1036)
    assert np.abs(Em - Em.transpose(0, 2, 1)).max() <= 1e-13 * np.abs(Em).max()
    assert np.abs(Bm - Bm.transpose(0, 2, 1)).max() > 1e-4 * np.abs(Bm).max()


def test_full_tensor_matinv_quirk():
    g, t = load_golden("wg_tiny")
    er, ur = np.moveaxis(g["full_er"], 0, 2), np.moveaxis(g["full_ur"], 0, 2)
    Em, Bm = O.element_matrices(t.nodes, t.tets, t.edges, t.tris, t.edge_lengths, t.tet_to_field, t.tet_to_edge,
                                ur, er, np.arange(4))
    assert np.abs(Em - g["full_elemE"]).max() <= RTOL * np.abs(g["full_elemE"]).max()
    assert np.abs(Bm - g["full_elemB"]).max() <= RTOL * np.abs(g["full_elemB"]).max()
    s = ur[:, :, :1]
    inv = O.matinv(s)[:, :, 0]
    det = np.linalg.det(s[:, :, 0])
    assert np.allclose(inv @ s[:, :, 0], det ** 2 * np.eye(3))          # adj*det, not adj/det


@pytest.mark.parametrize("name", ["wg_tiny", "wg_materials", "abc_lumped"])
def test_csr_pattern_and_values(name):
    g, t = load_golden(name)
    E, Bm = O.assemble_EB(t.nodes, t.tets, t.edges, t.tris, t.edge_lengths, t.tet_to_field, t.tet_to_edge, g["ur"], g["er"])
    assert np.array_equal(E.indptr, g["E_indptr"]) and np.array_equal(E.indices, g["E_indices"])   # bit-exact pattern
    assert np.abs(E.data - g["E_data"]).max() <= RTOL * np.abs(g["E_data"]).max()
    assert np.abs(Bm.data - g["B_data"]).max() <= RTOL * np.abs(g["B_data"]).max()


def _surface_terms(g, t, k0):
    N = t.n_field
    if "E_indptr" in g:
        E, Bm = csr(g, "E", N), csr(g, "E", N, "B_data")
    else:       # checksum-only fixture: E, B from the oracle (pinned by the E_dot_v / B_dot_v checks below)
        E, Bm = O.assemble_EB(t.nodes, t.tets, t.edges, t.tris, t.edge_lengths, t.tet_to_field, t.tet_to_edge, g["ur"], g["er"])
        v = g["probe_v"]
        assert np.abs(E @ v - g["E_dot_v"]).max() <= 1e-11 * np.abs(g["E_dot_v"]).max()
        assert np.abs(Bm @ v - g["B_dot_v"]).max() <= 1e-11 * np.abs(g["B_dot_v"]).max()
    return oracle_system(g, t, k0, E, Bm)


@pytest.mark.parametrize("name", ["wg_tiny", "wg_materials", "abc_lumped", "modal_microstrip", "lossy_slabs"])
def test_surface_terms_and_Kf(name):
    g, t = load_golden(name)
    k0 = 2 * np.pi * g["freqs"][0] / 299792458
    K, bvecs = _surface_terms(g, t, k0)
    ref = g["K_dot_v_0"]
    assert np.abs(K @ g["probe_v"] - ref).max() <= 1e-11 * np.abs(ref).max()
    for pn, bv in bvecs.items():
        r = g[f"bvec_0_p{pn}"]
        assert np.abs(bv - r).max() <= 1e-11 * np.abs(r).max()


def test_reference_solution_residuals_are_tiny():
    g, _ = load_golden("wg_medium")
    for k in g:
        if k.startswith("xres_"):
            assert g[k] < 1e-10
