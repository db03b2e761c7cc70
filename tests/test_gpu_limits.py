"""GPU: the documented hard limits of the library fail loudly (status + message) instead of corrupting memory:
more than 102 tetrahedra around one dof (emb_symbolic), surface ids beyond 16, lockstep groups wider than 4."""
import numpy as np
import pytest

from emerge_b200.lib import EmergeB200Error
from emerge_b200.synthmesh import mesh_tables

pytestmark = pytest.mark.gpu


def _orange(n):
    """n tetrahedra around one common edge (a fan of 'orange slices')"""
    ang = np.linspace(0, 2 * np.pi, n, endpoint=False)
    nodes = np.vstack([[0, 0, -1.0], [0, 0, 1.0], np.stack([np.cos(ang), np.sin(ang), 0 * ang], axis=1)])
    tets = np.array([[0, 1, 2 + i, 2 + (i + 1) % n] for i in range(n)], dtype=np.int64)
    p = nodes[tets]
    det = np.einsum("ij,ij->i", np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]), p[:, 3] - p[:, 0])
    tets[det < 0] = tets[det < 0][:, [0, 2, 1, 3]]
    return mesh_tables(nodes, tets)


def test_too_many_tets_around_a_dof_is_reported(gpu_ctx):
    ok = _orange(90)
    gpu_ctx.upload_mesh(ok.nodes, ok.tets, ok.tris, ok.tet_to_field, ok.tri_to_field, ok.edges.shape[1])
    gpu_ctx.symbolic()                                   # 90 tetrahedra share the axis edge: within the limit
    assert gpu_ctx.nnz > 0
    bad = _orange(130)
    gpu_ctx.upload_mesh(bad.nodes, bad.tets, bad.tris, bad.tet_to_field, bad.tri_to_field, bad.edges.shape[1])
    with pytest.raises(EmergeB200Error, match="shared by more than"):
        gpu_ctx.symbolic()


def test_surface_and_lockstep_limits(gpu_ctx):
    from tests.util import load_golden, golden_bcs
    from emerge_b200.sweep import FrequencySweep
    g, t = load_golden("wg_tiny")
    sw = FrequencySweep(t, g["er"], g["ur"], golden_bcs(g, t), recycle=0)
    sw.f_ref = 9e9
    sw.setup()
    sw.assemble_frequency(9e9)
    with pytest.raises(EmergeB200Error):
        sw.ctx.surface_define(16, np.array([0, 1]), frame=1)                  # sids 0..15
    sid = sw.sid[id(sw.ports[0])]
    with pytest.raises(EmergeB200Error):
        sw.ctx.solve_multi([sid] * 5, want_x=False)                           # lockstep groups of at most 4
    xs, infos = sw.ctx.solve_multi([sid] * 3, want_x=False, rtol=1e-9)       # three: padded with a zero column
    assert all(i["converged"] for i in infos)
    sw.ctx.close()
