"""CPU: the product's element templates (emerge_b200/csrc/ned2_tet.cuh) instantiated on the host, against the
reference's element matrices.  Checks the canonical-vertex formulation without a GPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from tests.util import load_golden

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hostlib():
    src = os.path.join(HERE, "hostcheck", "ned2_host.cpp")
    so = os.path.join(HERE, "hostcheck", "ned2_host.so")
    hdrs = [os.path.join(HERE, "..", "emerge_b200", "csrc", h) for h in ("ned2_tet.cuh", "ned2_fused.cuh")]
    if not os.path.exists(so) or os.path.getmtime(so) < max([os.path.getmtime(src)] + [os.path.getmtime(h) for h in hdrs]):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
    return ctypes.CDLL(so)


def _run(lib, p, vid, ur, er, fn="ned2_host_element"):
    K = np.zeros((20, 20), complex)
    M = np.zeros((20, 20), complex)
    arrs = [np.ascontiguousarray(p, dtype=float), np.ascontiguousarray(vid, dtype=np.int64),
            np.ascontiguousarray(ur, dtype=complex), np.ascontiguousarray(er, dtype=complex), K, M]
    getattr(lib, fn)(*[a.ctypes.data_as(ctypes.c_void_p) for a in arrs])
    return K, M


def test_host_templates_match_reference(hostlib):
    g, t = load_golden("wg_tiny")
    for it in range(6):
        K, M = _run(hostlib, t.nodes[:, t.tets[:, it]].T, t.tets[:, it], g["ur"][:, :, it], g["er"][:, :, it])
        assert np.abs(K - g["elemE"][it]).max() <= 1e-12 * np.abs(g["elemE"][it]).max()
        assert np.abs(M - g["elemB"][it]).max() <= 1e-12 * np.abs(g["elemB"][it]).max()
    for it in range(4):
        K, M = _run(hostlib, t.nodes[:, t.tets[:, it]].T, t.tets[:, it], g["full_ur"][it], g["full_er"][it])
        assert np.abs(K - g["full_elemE"][it]).max() <= 1e-12 * np.abs(g["full_elemE"][it]).max()
        assert np.abs(M - g["full_elemB"][it]).max() <= 1e-12 * np.abs(g["full_elemB"][it]).max()


def test_invariance_under_vertex_relabelling(hostlib):
    """The element matrix in reference-local order must not depend on the global ids' order."""
    g, t = load_golden("wg_tiny")
    rng = np.random.default_rng(0)
    p = t.nodes[:, t.tets[:, 3]].T
    ur, er = g["ur"][:, :, 3], g["er"][:, :, 3]
    K0, M0 = _run(hostlib, p, np.array([10, 20, 30, 40]), ur, er)
    assert np.abs(K0 - K0.T).max() < 1e-13 * np.abs(K0).max()
    # rank 11 = 20 - 9 gradients
    assert np.linalg.matrix_rank(K0, tol=1e-9 * np.abs(K0).max()) == 11


def test_fused_row_form_matches_templates_and_reference(hostlib):
    """ned2_fused.cuh (table-driven rows of the fused assembly kernel) against the compile-time templates and against the
    reference's element matrices, including full non-symmetric tensors (mirror rule, index typo, matinv quirk)."""
    g, t = load_golden("wg_tiny")
    rng = np.random.default_rng(5)
    cases = [(t.nodes[:, t.tets[:, it]].T, t.tets[:, it], g["ur"][:, :, it], g["er"][:, :, it], g["elemE"][it], g["elemB"][it])
             for it in range(6)]
    cases += [(t.nodes[:, t.tets[:, it]].T, t.tets[:, it], g["full_ur"][it], g["full_er"][it], g["full_elemE"][it],
               g["full_elemB"][it]) for it in range(4)]
    for p, vid, ur, er, Eref, Bref in cases:
        for perm in (np.arange(4), rng.permutation(4), rng.permutation(4)):
            # relabelled global ids exercise every canonical ordering; the reference-order result must not change
            ids = np.array([10, 20, 30, 40])[perm] if perm is not None else vid
            K0, M0 = _run(hostlib, p, ids, ur, er)
            for fn in ("ned2_host_fused", "ned2_host_fused_pairs", "ned2_host_fused_flat"):
                K1, M1 = _run(hostlib, p, ids, ur, er, fn=fn)
                assert np.abs(K1 - K0).max() <= 1e-14 * np.abs(K0).max()
                assert np.abs(M1 - M0).max() <= 1e-14 * np.abs(M0).max()
        K1, M1 = _run(hostlib, p, vid, ur, er, fn="ned2_host_fused_flat")
        assert np.abs(K1 - Eref).max() <= 1e-12 * np.abs(Eref).max()
        assert np.abs(M1 - Bref).max() <= 1e-12 * np.abs(Bref).max()
