"""GPU: element matrices of the port boundary-mode analysis (emb_bma_element_matrices, csrc/bma.cuh) against the unmodified
reference (tests/golden/bma_microstrip.npz: generalized_matrix_GQ with full tensors, assembled E / B of
Assembler.assemble_bma_matrices) and, on a larger jittered port mesh with random lossy tensors, against the oracle."""
import numpy as np
import pytest

from oracle import ned2_oracle as O
from tests.test_host_bma import load_bma, assemble, ref_csr

pytestmark = pytest.mark.gpu


def test_reference_element_and_assembled_matrices(gpu_ctx):
    g = load_bma()
    k0 = float(g["k0"])
    eA, eB = gpu_ctx.bma_element_matrices(g["xy"], g["s_tris"], g["s_edges"], g["s_tri_to_edge"], g["er"], g["ur"], k0)
    E, B = assemble(g, eA, eB)
    for M, name in ((E, "E"), (B, "B")):
        R = ref_csr(g, name)
        assert abs(M - R).max() <= 1e-12 * abs(R).max(), name
    sel = g["full_sel"]
    fA, fB = gpu_ctx.bma_element_matrices(g["xy"], g["s_tris"][:, sel], g["s_edges"], g["s_tri_to_edge"][:, sel],
                                          np.moveaxis(g["full_er"], 0, 2), np.moveaxis(g["full_ur"], 0, 2), k0)
    assert np.abs(fA - g["full_A"]).max() <= 1e-12 * np.abs(g["full_A"]).max()
    assert np.abs(fB - g["full_B"]).max() <= 1e-12 * np.abs(g["full_B"]).max()


def test_larger_port_mesh_against_oracle(gpu_ctx):
    rng = np.random.default_rng(5)
    nx, ny = 60, 40
    X, Y = np.meshgrid(np.linspace(0, 30e-3, nx + 1), np.linspace(0, 20e-3, ny + 1), indexing="ij")
    xy = np.stack([X.ravel(), Y.ravel()])
    xy += (rng.random(xy.shape) - 0.5) * 0.2e-3
    nid = lambda i, j: i * (ny + 1) + j
    I, J = [a.ravel() for a in np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")]
    t1 = np.stack([nid(I, J), nid(I + 1, J), nid(I + 1, J + 1)])
    t2 = np.stack([nid(I, J), nid(I, J + 1), nid(I + 1, J + 1)])
    tris = np.sort(np.concatenate([t1, t2], axis=1), axis=0)
    # random vertex relabelling so the three local edge directions vary, then ascending vertex order per triangle
    perm = rng.permutation(xy.shape[1])
    xy = xy[:, np.argsort(perm)]
    tris = np.sort(perm[tris], axis=0)
    ek = np.concatenate([tris[[0, 1]], tris[[1, 2]], tris[[0, 2]]], axis=1)
    edges, inv = np.unique(ek, axis=1, return_inverse=True)
    nt = tris.shape[1]
    t2e = inv.reshape(3, nt)
    flip = rng.random(edges.shape[1]) < 0.5                      # edge directions independent of the vertex order
    edges[:, flip] = edges[::-1, flip]
    er = rng.standard_normal((3, 3, nt)) * 0.2 + 1j * rng.standard_normal((3, 3, nt)) * 0.1 + (4 - 0.1j) * np.eye(3)[:, :, None]
    ur = np.repeat(np.diag([1.5, 1.2 - 0.05j, 1.0]).astype(complex)[:, :, None], nt, axis=2)
    ur[:, :, ::3] += rng.standard_normal((3, 3, len(range(0, nt, 3)))) * 0.1            # some full tensors: matinv quirk
    eA, eB = gpu_ctx.bma_element_matrices(xy, tris, edges, t2e, er, ur, 83.7)
    rA, rB = O.bma_element_matrices(xy, tris, edges, t2e, er, ur, 83.7)
    assert np.abs(eA - rA).max() <= 1e-11 * np.abs(rA).max()
    assert np.abs(eB - rB).max() <= 1e-11 * np.abs(rB).max()
    eA2, eB2 = gpu_ctx.bma_element_matrices(xy, tris, edges, t2e, er, ur, 83.7)
    assert np.array_equal(eA, eA2) and np.array_equal(eB, eB2)
    with pytest.raises(Exception):
        bad = t2e.copy()
        bad[0, 0] = (bad[0, 0] + 7) % edges.shape[1]              # an edge that does not belong to its triangle
        gpu_ctx.bma_element_matrices(xy, tris, edges, bad, er, ur, 83.7)


def test_shift_invert_operator_and_eigen_solve(gpu_ctx):
    """emb_shift_invert_*: the device operator equals (A - sigma B)^-1 B of numpy on a random complex pair (pivoting
    exercised by a zero diagonal), and modal.gpu_eig on the reference port matrices returns the reference's mode."""
    from emerge_b200 import modal
    from tests.test_host_bma import reference_target
    rng = np.random.default_rng(2)
    n = 333
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    A[np.arange(n), np.arange(n)] = 0.0
    B = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    sigma = 0.3 - 0.2j
    gpu_ctx.shift_invert_setup(A, B, sigma)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    y = gpu_ctx.shift_invert_apply(x)
    ref = np.linalg.solve(A - sigma * B, B @ x)
    assert np.linalg.norm(y - ref) <= 1e-9 * np.linalg.norm(ref)
    gpu_ctx.shift_invert_free()
    with pytest.raises(Exception):
        gpu_ctx.shift_invert_apply(x)
    g = load_bma()
    lam, V = modal.gpu_eig(gpu_ctx, ref_csr(g, "E"), ref_csr(g, "B"), g["solve_ids"], 1, True, reference_target(g))
    assert abs(np.sqrt(-lam[0]).real - float(g["beta"])) <= 1e-9 * float(g["beta"])
