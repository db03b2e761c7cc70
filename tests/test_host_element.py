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
    hdr = os.path.join(HERE, "..", "emerge_b200", "csrc", "ned2_tet.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
    return ctypes.CDLL(so)


def _run(lib, p, vid, ur, er):
    K = np.zeros((20, 20), complex)
    M = np.zeros((20, 20), complex)
    arrs = [np.ascontiguousarray(p, dtype=float), np.ascontiguousarray(vid, dtype=np.int64),
            np.ascontiguousarray(ur, dtype=complex), np.ascontiguousarray(er, dtype=complex), K, M]
    lib.ned2_host_element(*[a.ctypes.data_as(ctypes.c_void_p) for a in arrs])
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
