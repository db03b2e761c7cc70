"""CPU: the product's boundary-mode element matrices (emerge_b200/csrc/bma.cuh) instantiated on the host, against the
unmodified reference (tests/golden/bma_microstrip.npz): element matrices of generalized_matrix_GQ with full non-symmetric
tensors, and the assembled E, B of Assembler.assemble_bma_matrices (assembler.py:246-308)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

from tests.util import GOLDEN

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hostlib():
    src = os.path.join(HERE, "hostcheck", "bma_host.cpp")
    so = os.path.join(HERE, "hostcheck", "bma_host.so")
    hdrs = [os.path.join(HERE, "..", "emerge_b200", "csrc", h) for h in ("bma.cuh", "ned2_tet.cuh", "emb_common.cuh")]
    if not os.path.exists(so) or os.path.getmtime(so) < max([os.path.getmtime(src)] + [os.path.getmtime(h) for h in hdrs]):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
    lib = ctypes.CDLL(so)
    lib.bma_host_element.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]
    return lib


def load_bma():
    return dict(np.load(os.path.join(GOLDEN, "bma_microstrip.npz"), allow_pickle=False))


def local_map(g, it):
    """local vertices of the triangle's three edges in the direction of the global edge (local_tri_to_edgeid)"""
    v = list(g["s_tris"][:, it])
    return np.array([[v.index(g["s_edges"][0, e]), v.index(g["s_edges"][1, e])] for e in g["s_tri_to_edge"][:, it]], dtype=np.int32)


def element(lib, g, it, ur, er):
    A = np.zeros((14, 14), complex)
    B = np.zeros((14, 14), complex)
    xy = np.ascontiguousarray(g["xy"][:, g["s_tris"][:, it]].T, dtype=float)
    args = [xy, local_map(g, it), np.ascontiguousarray(ur, dtype=complex), np.ascontiguousarray(er, dtype=complex)]
    lib.bma_host_element(*[a.ctypes.data_as(ctypes.c_void_p) for a in args], float(g["k0"]),
                         A.ctypes.data_as(ctypes.c_void_p), B.ctypes.data_as(ctypes.c_void_p))
    return A, B


def assemble(g, eA, eB):
    """COO -> CSR exactly as generelized_eigenvalue_matrix does (nedeleclegrange2.py:47-52)"""
    ttf = g["tri_to_field"]
    nt, n = ttf.shape[1], int(g["n_field"])
    rows = np.repeat(ttf.T[:, :, None], 14, axis=2).ravel()
    cols = np.repeat(ttf.T[:, None, :], 14, axis=1).ravel()
    E = sp.coo_matrix((eA.ravel(), (rows, cols)), shape=(n, n)).tocsr()
    B = sp.coo_matrix((eB.ravel(), (rows, cols)), shape=(n, n)).tocsr()
    return E, B


def ref_csr(g, name):
    n = int(g["n_field"])
    return sp.csr_matrix((g[name + "_data"], g[name + "_indices"], g[name + "_indptr"]), shape=(n, n))


def test_element_matrices_with_full_tensors(hostlib):
    g = load_bma()
    for k, it in enumerate(g["full_sel"]):
        A, B = element(hostlib, g, int(it), g["full_ur"][k], g["full_er"][k])
        assert np.abs(A - g["full_A"][k]).max() <= 1e-12 * np.abs(g["full_A"][k]).max()
        assert np.abs(B - g["full_B"][k]).max() <= 1e-12 * np.abs(g["full_B"][k]).max()
        assert np.abs(g["full_B"][k] - g["full_B"][k].T).max() > 1e-3 * np.abs(g["full_B"][k]).max()   # genuinely non-symmetric


def test_assembled_matrices_match_reference(hostlib):
    g = load_bma()
    nt = g["s_tris"].shape[1]
    eA = np.zeros((nt, 14, 14), complex)
    eB = np.zeros((nt, 14, 14), complex)
    for it in range(nt):
        eA[it], eB[it] = element(hostlib, g, it, g["ur"][:, :, it], g["er"][:, :, it])
    E, B = assemble(g, eA, eB)
    for M, name in ((E, "E"), (B, "B")):
        R = ref_csr(g, name)
        assert abs(M - R).max() <= 1e-12 * abs(R).max(), name


def test_oracle_restatement_matches_reference():
    """oracle.bma_element_matrices (numpy, vectorised over the triangles) against the same reference outputs"""
    from oracle import ned2_oracle as O
    g = load_bma()
    eA, eB = O.bma_element_matrices(g["xy"], g["s_tris"], g["s_edges"], g["s_tri_to_edge"], g["er"], g["ur"], float(g["k0"]))
    E, B = assemble(g, eA, eB)
    for M, name in ((E, "E"), (B, "B")):
        R = ref_csr(g, name)
        assert abs(M - R).max() <= 1e-12 * abs(R).max(), name
    sel = g["full_sel"]
    fer = np.moveaxis(g["full_er"], 0, 2)
    fur = np.moveaxis(g["full_ur"], 0, 2)
    eA, eB = O.bma_element_matrices(g["xy"], g["s_tris"][:, sel], g["s_edges"], g["s_tri_to_edge"][:, sel], fer, fur, float(g["k0"]))
    assert np.abs(eA - g["full_A"]).max() <= 1e-12 * np.abs(g["full_A"]).max()
    assert np.abs(eB - g["full_B"]).max() <= 1e-12 * np.abs(g["full_B"]).max()


class _HostShiftInvert:
    """numpy stand-in for the device operator (emb_shift_invert_*), so the eigen-solve logic can be pinned without a GPU"""

    def shift_invert_setup(self, A, B, sigma):
        self.M, self.B = np.linalg.inv(A - sigma * B), B

    def shift_invert_apply(self, x):
        return self.M @ (self.B @ x)

    def shift_invert_free(self):
        pass


def reference_target(g):
    """target_kz of modal_analysis(TEM=True) (emfreq3d.py:270-272): mean(er) * mean(ur) * 1.1 * k0"""
    er, ur = g["er"], g["ur"]
    return np.mean(er[er > 0]) * np.mean(ur[ur > 0]) * 1.1 * float(g["k0"])


def test_eigen_solve_logic_finds_the_reference_mode():
    """modal.gpu_eig (shift-invert Arnoldi + filter_real_modes) returns the propagation constant the reference's dense
    LAPACK path selected for this port (mode.beta of the fixture)."""
    from emerge_b200 import modal
    g = load_bma()
    lam, V = modal.gpu_eig(_HostShiftInvert(), ref_csr(g, "E"), ref_csr(g, "B"), g["solve_ids"], 1, True, reference_target(g))
    assert len(lam) >= 1 and V.shape == (len(g["solve_ids"]), len(lam))
    assert abs(np.sqrt(-lam[0]).real - float(g["beta"])) <= 1e-9 * float(g["beta"])
    ids = g["solve_ids"]
    ix = np.ix_(ids, ids)
    r = ref_csr(g, "E")[ix] @ V[:, 0] - lam[0] * (ref_csr(g, "B")[ix] @ V[:, 0])
    assert np.linalg.norm(r) <= 1e-6 * np.linalg.norm(ref_csr(g, "E")[ix] @ V[:, 0])        # eigen-residual
