"""Eigen-solve of the port boundary-mode analysis with the operator on the device.

Mirror of SolveRoutine.eig (reference fem/solver.py:471-505) and of the two solvers behind it:
  SolverLAPACK.eig  (:311-342)  dense scipy.linalg.eig (QZ on one core, O(25 n^3)) + filter_real_modes (:34-68)
  SolverARPACK.eig  (:344-357)  eigsh(A, k, M=B, sigma=-target_kz^2): ARPACK around a SuperLU factorisation
Both look for the modes whose propagation constant is closest to an estimate.  Here the spectral transformation
    v -> (A - sigma B)^-1 B v,   sigma = -target_kz^2
is a device-resident dense operator (emb_shift_invert_setup: Gauss-Jordan inversion with partial pivoting, csrc/modal.cu) and
ARPACK's implicitly restarted Arnoldi iteration (scipy.sparse.linalg.eigs - vectors of the port size, host) drives it.
Return values follow the reference: direct=True -> filter_real_modes' selection and ordering (closest beta first),
direct=False -> the `nmodes` eigenvalues nearest sigma in ascending order, as eigsh returns them.
"""
from __future__ import annotations

import types

import numpy as np


def filter_real_modes(eigvals, eigvecs, k0, ermax=2, urmax=2):
    """fem/solver.py:34-68: keep -2 k0^2 ermax urmax <= lambda <= -1 and order by |sqrt(-lambda) - k0|"""
    upper = -(k0 ** 2) * ermax * urmax * 2
    mask = (eigvals <= -1) & (eigvals >= upper)
    vals, vecs = eigvals[mask], eigvecs[:, mask]
    order = np.argsort(np.abs(np.sqrt(-vals) - k0))
    return vals[order], vecs[:, order]


def gpu_eig(ctx, A, B, solve_ids, nmodes: int = 6, direct=None, target_k0=None, which: str = "LM", tol: float = 1e-12):
    """(eigen_values, eigen_modes) of A x = lambda B x restricted to solve_ids, the modes nearest -target_k0^2.
    Same arguments as SolveRoutine.eig (fem/solver.py:471-505)."""
    from scipy.sparse.linalg import LinearOperator, eigs
    solve_ids = np.asarray(solve_ids)
    ix = np.ix_(solve_ids, solve_ids)
    Ad = np.asarray(A[ix].toarray() if hasattr(A, "toarray") else np.asarray(A)[ix], dtype=np.complex128)
    Bd = np.asarray(B[ix].toarray() if hasattr(B, "toarray") else np.asarray(B)[ix], dtype=np.complex128)
    n = Ad.shape[0]
    if target_k0 is None:
        target_k0 = 0.0
    sigma = -(complex(target_k0) ** 2)
    k = int(max(1, min(max(nmodes, 6) if direct else nmodes, n - 2)))
    ctx.shift_invert_setup(Ad, Bd, sigma)
    try:
        op = LinearOperator((n, n), matvec=ctx.shift_invert_apply, dtype=np.complex128)
        v0 = np.random.default_rng(0).standard_normal(n) + 0j                    # deterministic start vector
        theta, V = eigs(op, k=k, which="LM", v0=v0, tol=tol, ncv=min(n - 1, max(4 * k + 8, 40)))
    finally:
        ctx.shift_invert_free()
    if not np.all(np.isfinite(theta)):
        raise RuntimeError("gpu_eig: A - sigma B is singular to working precision (shift equals an eigenvalue)")
    lam = sigma + 1.0 / theta
    if np.abs(lam.imag).max() <= 1e-9 * max(np.abs(lam.real).max(), 1e-300) and abs(sigma.imag) == 0:
        lam = lam.real                      # lossless port: real spectrum, as the reference's eigsh / filtered eig return it
    if direct or direct is None:
        return filter_real_modes(lam, V, complex(target_k0).real if np.isrealobj(lam) else target_k0)
    order = np.argsort(lam.real)
    return lam[order], V[:, order]


def install_modal(physics, asm):
    """Electrodynamics3D.modal_analysis (emfreq3d.py:201-364) then runs its element loop (GpuAssembler.assemble_bma_matrices)
    and the spectral transformation of its eigen-solve on the device: `physics.solveroutine.eig` is replaced."""
    routine = physics.solveroutine

    def eig(self, A, B, solve_ids, nmodes: int = 6, direct=None, target_k0=None, which: str = "LM"):
        return gpu_eig(asm.ctx, A, B, solve_ids, nmodes, direct, target_k0, which)
    routine.eig = types.MethodType(eig, routine)
    return routine
