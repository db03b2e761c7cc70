"""Drop-in seams (emerge_b200/dropin.py): the Assembler / SolveRoutine interface of the reference on top of the C ABI.
GPU part: the golden waveguide driven exactly as emfreq3d.py:682-699 drives its assembler and solve routine.
CPU part: tag propagation of the lazy host objects (no compute)."""
import types

import numpy as np
import pytest

from emerge_b200.dropin import DeviceRHS, _ZeroRHS
from tests.util import load_golden, golden_bcs, csr, db_deg_close


def test_device_rhs_tag_survives_only_b_plus_port_vector():
    b = _ZeroRHS(5)
    pv = DeviceRHS(np.arange(5) + 0j, sid=3, owner="me")
    s = b + pv                                    # emfreq3d.py:691
    assert isinstance(s, DeviceRHS) and s._emb_sid == 3 and np.array_equal(np.asarray(s), np.arange(5))
    assert getattr(pv + b, "_emb_sid", None) == 3
    assert getattr(pv * 2, "_emb_sid", None) is None
    assert getattr(pv + pv, "_emb_sid", None) is None
    b2 = _ZeroRHS(5)
    b2[1] = 1.0                                   # no longer zero: the sum must be solved from its host values
    assert getattr(b2 + pv, "_emb_sid", None) is None
    assert not getattr(b + b, "_emb_zero", False)


def test_device_rhs_inplace_arithmetic_and_frequency_tag():
    pv = DeviceRHS(np.arange(5) + 0j, sid=3, owner="me", freq=9e9)
    assert (_ZeroRHS(5) + pv)._emb_freq == 9e9
    pv += np.ones(5)                              # used to recurse forever (out= forwarded to __array_ufunc__)
    assert np.array_equal(np.asarray(pv), np.arange(5) + 1) and pv._emb_sid is None
    q = DeviceRHS(np.zeros(3, complex), sid=1, owner="me", freq=1.0)
    np.multiply(q, 2.0, out=q)
    assert q._emb_sid is None


class _Field:
    """what GpuAssembler reads from a reference Nedelec2: .mesh (+get_triangles) and the dof tables"""

    def __init__(self, t):
        self.mesh = types.SimpleNamespace(**{k: getattr(t, k) for k in (
            "nodes", "tets", "edges", "tris", "tet_to_edge", "tet_to_tri", "tri_to_edge", "tri_to_tet", "edge_lengths")})
        self.tet_to_field, self.tri_to_field, self.edge_to_field = t.tet_to_field, t.tri_to_field, t.edge_to_field


@pytest.mark.gpu
def test_assembler_and_solver_seams_reproduce_reference_outputs():
    from emerge_b200.dropin import GpuAssembler, DeviceCSR
    g, t = load_golden("wg_tiny")
    bcs = golden_bcs(g, t)
    ports = bcs[1:]
    asm = GpuAssembler(rtol=1e-11)
    field = _Field(t)
    N = t.n_field
    S = np.zeros((len(g["freqs"]), 2, 2), dtype=complex)
    for i, f in enumerate(g["freqs"]):
        K, b, solve_ids, pv = asm.assemble_freq_matrix(field, g["er"], g["ur"], bcs, f, cache_matrices=True)
        assert isinstance(K, DeviceCSR) and K.shape == (N, N) and b.shape == (N,) and not b.any()
        assert np.array_equal(solve_ids, g["solve_ids"])
        for p in ports:
            ref = g[f"bvec_{i}_p{p.port_number}"]
            assert np.abs(np.asarray(pv[p.port_number]) - ref).max() <= 1e-11 * np.abs(ref).max()
        if i == 0:       # the lazily downloaded K equals the reference's K on the solve space
            Kh = K.materialize()
            s = g["solve_ids"]
            Kr = csr(g, "K0", N)[s][:, s].tocsr()
            d = (Kh[s][:, s] - Kr).tocsr()
            assert np.abs(d.data).max() <= 1e-11 * np.abs(Kr.data).max()
        for p in ports:  # emfreq3d.py:688-699
            b_active = b + pv[p.port_number]
            assert getattr(b_active, "_emb_sid", None) is not None
            x = asm.solve(K, b_active, solve_ids, reuse=True)
            xr = g[f"x_{i}_p{p.port_number}"]
            assert x.shape == (N,) and np.linalg.norm(x - xr) <= 1e-7 * np.linalg.norm(xr)
            # same system through a plain host vector (upload path)
            x2 = asm.solve(K, np.asarray(pv[p.port_number]).copy(), solve_ids)
            assert np.linalg.norm(x2 - xr) <= 1e-7 * np.linalg.norm(xr)
    with pytest.raises(Exception):
        import scipy.sparse as sp
        asm.solve(sp.identity(N, format="csr", dtype=complex), np.zeros(N, complex), g["solve_ids"])
    asm.ctx.close()
