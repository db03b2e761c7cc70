"""Host logic of the reduced-basis projection: the small complex least-squares solver (csrc/lsq.hpp, Householder QR)
against numpy.linalg.lstsq, including a rank-deficient basis and more columns than the recycling ever holds."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostcheck", "lsq_host.cpp")
SO = os.path.join(HERE, "hostcheck", "lsq_host.so")


@pytest.fixture(scope="module")
def lsq():
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", SO, SRC])
    lib = C.CDLL(SO)
    lib.lsq_solve.argtypes = [C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 4
    lib.lsq_solve.restype = None
    lib.herm_eig.argtypes = [C.c_int] + [C.c_void_p] * 3
    lib.herm_eig.restype = None

    def solve(G, g):
        nq, m = G.shape
        nv = g.shape[1]
        Gc = np.asfortranarray(G.astype(np.complex128))
        gc = np.asfortranarray(g.astype(np.complex128))
        y = np.zeros((nv, m), dtype=np.complex128)             # routine layout: y[j * nv + k]
        y = np.zeros((m, nv), dtype=np.complex128)
        res = np.zeros(nv)
        lib.lsq_solve(nq, m, nv, Gc.ctypes.data, gc.ctypes.data, y.ctypes.data, res.ctypes.data)
        return y, res
    solve.lib = lib
    return solve


@pytest.mark.parametrize("nq,m,nv", [(8, 3, 1), (88, 22, 2), (160, 40, 4), (5, 5, 2)])
def test_matches_numpy_lstsq(lsq, nq, m, nv):
    rng = np.random.default_rng(nq * 100 + m)
    G = rng.standard_normal((nq, m)) + 1j * rng.standard_normal((nq, m))
    G *= np.logspace(0, -6, m)[None, :]                        # badly scaled columns, as the recycled directions are
    g = rng.standard_normal((nq, nv)) + 1j * rng.standard_normal((nq, nv))
    y, res = lsq(G, g)
    yr, *_ = np.linalg.lstsq(G, g, rcond=None)
    rr = np.linalg.norm(g - G @ yr, axis=0)
    assert np.allclose(G @ y, G @ yr, rtol=0, atol=1e-10 * np.linalg.norm(g))
    assert np.allclose(res, rr, rtol=1e-8, atol=1e-12 * np.linalg.norm(g))


def test_rank_deficient_basis_is_harmless(lsq):
    rng = np.random.default_rng(7)
    G = rng.standard_normal((30, 6)) + 1j * rng.standard_normal((30, 6))
    G[:, 4] = G[:, 1] * (2 - 1j)                               # a direction already inside the span
    g = (G[:, :4] @ (rng.standard_normal((4, 2)) + 0j))        # right-hand sides inside the span
    y, res = lsq(G, g)
    assert np.all(np.isfinite(y))
    assert np.linalg.norm(g - G @ y) <= 1e-10 * np.linalg.norm(g)
    assert np.all(res <= 1e-10 * np.linalg.norm(g))


@pytest.mark.parametrize("n", [1, 2, 7, 40])
def test_hermitian_jacobi_eigensolver(lsq, n):
    """Gram matrices of the reduced basis are Hermitian and nearly singular: eigenvalues over 16 decades."""
    rng = np.random.default_rng(n)
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    lam_true = np.logspace(0, -16, n) if n > 1 else np.array([2.5])
    H = (Q * lam_true) @ Q.conj().T
    H = 0.5 * (H + H.conj().T)
    Hc = np.ascontiguousarray(H, dtype=np.complex128)
    V = np.zeros((n, n), dtype=np.complex128)
    lam = np.zeros(n)
    lsq.lib.herm_eig(n, Hc.ctypes.data, V.ctypes.data, lam.ctypes.data)
    assert np.allclose(V.conj().T @ V, np.eye(n), atol=1e-12)                       # unitary
    assert np.allclose((V * lam) @ V.conj().T, H, atol=1e-13 * np.abs(H).max())     # reconstructs H
    assert np.allclose(np.sort(lam)[::-1][: max(1, n // 2)], np.sort(lam_true)[::-1][: max(1, n // 2)], rtol=1e-8)
