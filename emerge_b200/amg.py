"""Smoothed-aggregation AMG setup for the nodal auxiliary problems of the multilevel preconditioner (host side, numpy /
scipy, one-time per mesh).  Only the SETUP lives here (aggregates, prolongators, Galerkin products); the cycles run on
the GPU (emerge_b200/csrc/amg.cuh) from the level matrices this module produces.

The reference has no counterpart: it solves every A(f) with a sparse direct solver (fem/solver.py:243-309).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def _rowmax(indptr, vals, n, fill):
    """max of vals over each CSR row (empty rows -> fill)."""
    out = np.full(n, fill, dtype=vals.dtype)
    nz = np.diff(indptr) > 0
    if len(vals):
        out[nz] = np.maximum.reduceat(vals, indptr[:-1][nz])
    return out


def mis2_aggregate(S: sp.csr_matrix, seed: int = 0) -> np.ndarray:
    """Aggregates from a distance-2 maximal independent set of the (symmetric, diagonal-free) strength graph S.
    Vectorised (a handful of rounds of neighbour-maxima over the CSR arrays).  Returns agg[i] in [0, n_agg)."""
    n = S.shape[0]
    indptr, indices = S.indptr, S.indices
    rng = np.random.default_rng(seed)
    key = rng.permutation(n).astype(np.int64) + 1            # distinct positive keys
    state = np.zeros(n, dtype=np.int8)                        # 0 undecided, 1 root, -1 covered
    state[np.diff(indptr) == 0] = 1                           # isolated nodes are their own aggregates
    while True:
        und = state == 0
        if not und.any():
            break
        k = np.where(und, key, 0)
        m1 = np.maximum(k, _rowmax(indptr, k[indices], n, 0))
        m2 = np.maximum(m1, _rowmax(indptr, m1[indices], n, 0))
        root = und & (m2 == key)
        state[root] = 1
        r = (state == 1).astype(np.int8)
        c1 = np.maximum(r, _rowmax(indptr, r[indices], n, 0))
        c2 = np.maximum(c1, _rowmax(indptr, c1[indices], n, 0))
        state[(state == 0) & (c2 > 0)] = -1
    roots = np.nonzero(state == 1)[0]
    agg = np.full(n, -1, dtype=np.int64)
    agg[roots] = np.arange(len(roots))
    # distance-1 neighbours join their (unique by MIS-2) root; distance-2 nodes join any aggregated neighbour
    for _ in range(2):
        lab = np.where(agg >= 0, agg, -1)
        best = _rowmax(indptr, lab[indices], n, -1)
        take = (agg < 0) & (best >= 0)
        agg[take] = best[take]
    left = agg < 0
    if left.any():                                           # disconnected leftovers (should not happen)
        agg[left] = len(roots) + np.arange(left.sum())
    return agg


def strength(A: sp.csr_matrix, theta: float) -> sp.csr_matrix:
    """|a_ij| >= theta sqrt(a_ii a_jj), diagonal removed, symmetrised pattern."""
    A = A.tocoo()
    d = np.abs(A.tocsr().diagonal())
    d[d == 0] = 1.0
    keep = (A.row != A.col) & (np.abs(A.data) >= theta * np.sqrt(d[A.row] * d[A.col]))
    S = sp.coo_matrix((np.ones(keep.sum()), (A.row[keep], A.col[keep])), shape=A.shape).tocsr()
    S = ((S + S.T) > 0).astype(np.float64).tocsr()
    S.sort_indices()
    return S


def _rho_DinvA(A, dinv, iters=12, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(A.shape[0])
    lam = 1.0
    nrm = lambda v: float(np.sqrt(np.sum(v * v)))      # not BLAS: a threaded dot costs milliseconds on these sizes
    for _ in range(iters):
        y = dinv * (A @ x)
        ny = nrm(y)
        lam = ny / max(nrm(x), 1e-300)
        x = y / max(ny, 1e-300)
    return lam


def sa_hierarchy(A: sp.spmatrix, block: int = 1, theta: float = 0.06, coarse_size: int = 600, max_levels: int = 8,
                 omega: float = 4.0 / 3.0):
    """Smoothed-aggregation hierarchy of the real symmetric positive (semi-)definite matrix A.
    block > 1: the unknowns are `block` interleaved components per node (node-major: dof = node*block + comp); nodes are
    aggregated on the graph of summed |block| strengths and every component gets its own piecewise-constant column.
    Returns a list of levels [{A, P (to the next coarser level, or None), dinv, rho}], finest first."""
    A = sp.csr_matrix(A, dtype=np.float64)
    levels = []
    while True:
        A.sort_indices()
        d = A.diagonal().copy()
        d[d == 0] = 1.0
        dinv = 1.0 / d
        rho = _rho_DinvA(A, dinv)
        lev = dict(A=A, P=None, dinv=dinv, rho=rho)
        levels.append(lev)
        n = A.shape[0]
        if n <= coarse_size or len(levels) >= max_levels:
            break
        nn = n // block
        if block > 1:
            Ab = abs(A).tocoo()
            Ng = sp.coo_matrix((Ab.data, (Ab.row // block, Ab.col // block)), shape=(nn, nn)).tocsr()
        else:
            Ng = A
        agg = mis2_aggregate(strength(Ng, theta), seed=len(levels))
        na = int(agg.max()) + 1
        if na * block >= 0.8 * n:                               # no coarsening progress
            break
        rows = np.arange(n)
        cols = agg[rows // block] * block + rows % block
        T = sp.coo_matrix((np.ones(n), (rows, cols)), shape=(n, na * block)).tocsr()
        cnt = np.asarray(T.multiply(T).sum(axis=0)).ravel()
        T = T @ sp.diags(1.0 / np.sqrt(np.maximum(cnt, 1.0)))
        P = (T - sp.diags(omega / rho * dinv) @ (A @ T)).tocsr()
        P.sort_indices()
        lev["P"] = P
        A = (P.T @ A @ P).tocsr()
    return levels
