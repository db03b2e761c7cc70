"""CPU: sweep-progress checkpoint (resume without re-solving) and operator save / load round trip."""
import numpy as np
import pytest
import scipy.sparse as sp

from emerge_b200.checkpoint import SweepCheckpoint, load_operators, problem_fingerprint, save_operators
from tests.util import load_golden, golden_bcs


def test_resume_skips_solved_points_and_rejects_other_problems(tmp_path):
    g, t = load_golden("wg_tiny")
    bcs = golden_bcs(g, t)
    freqs = np.linspace(8e9, 12e9, 7)
    fp = problem_fingerprint(t, g["er"], g["ur"], bcs, freqs, 1e-8)
    assert fp == problem_fingerprint(t, g["er"], g["ur"], bcs, freqs, 1e-8)
    assert fp != problem_fingerprint(t, g["er"] * 2, g["ur"], bcs, freqs, 1e-8)
    assert fp != problem_fingerprint(t, g["er"], g["ur"], bcs, freqs[:-1], 1e-8)
    path = str(tmp_path / "sweep.npz")
    ck = SweepCheckpoint(path, fp, freqs, 2)
    order = [0, 6, 3, 1, 5, 2, 4]
    for i in order[:3]:                                  # the job "dies" after three points
        ck.on_point(i, np.full((2, 2), i + 1j), [dict(iters=3, relres=np.float64(1e-9), converged=np.bool_(True), freq=freqs[i])])
    ck2 = SweepCheckpoint(path, fp, freqs, 2)            # restart
    assert ck2.remaining(order) == order[3:]
    assert ck2.stats[6][0]["iters"] == 3 and ck2.stats[6][0]["converged"] is True
    S = ck2.merge_into(np.zeros((7, 2, 2), complex))
    assert np.all(S[[0, 6, 3]] != 0) and np.all(S[[1, 5, 2, 4]] == 0) and S[3, 0, 0] == 3 + 1j
    with pytest.raises(ValueError):
        SweepCheckpoint(path, "other", freqs, 2)
    with pytest.raises(ValueError):
        SweepCheckpoint(path, fp, freqs, 3)


def test_operator_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    A = sp.random(30, 30, 0.2, random_state=1, format="csr").astype(complex)
    E, B = A.copy(), A.copy()
    E.data = rng.standard_normal(A.nnz) + 1j * rng.standard_normal(A.nnz)
    B.data = rng.standard_normal(A.nnz) + 0j
    p = str(tmp_path / "ops.npz")
    save_operators(p, E, B)
    E2, B2 = load_operators(p)
    assert (E2 != E).nnz == 0 and (B2 != B).nnz == 0
    C = sp.identity(30, format="csr", dtype=complex)
    with pytest.raises(ValueError):
        save_operators(p, E, C)
