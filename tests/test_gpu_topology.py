"""GPU: mesh topology tables built on the device (csrc/topology.cu) against the host construction
(emerge_b200.synthmesh.mesh_tables, itself asserted equal to the reference's Mesh3D / Nedelec2 tables when the fixtures
were generated, tests/golden/make_golden.py::_tables_check).  Integer tables must be array-equal, edge lengths bit-equal."""
import numpy as np
import pytest

from emerge_b200.synthmesh import box_mesh, mesh_tables
from tests.util import load_golden

pytestmark = pytest.mark.gpu
NAMES = ("edges", "tris", "tet_to_edge", "tet_to_tri", "tri_to_edge", "tri_to_tet", "tet_to_field", "tri_to_field", "edge_to_field")


def _same(a, b):
    for n in NAMES:
        assert np.array_equal(getattr(a, n), getattr(b, n)), n
    assert np.array_equal(a.edge_lengths.view(np.int64), b.edge_lengths.view(np.int64))
    assert a.n_field == b.n_field


def test_lexicographic_numbering_equals_host_tables(gpu_ctx):
    box = box_mesh(7, 5, 9, 22.86e-3, 10.16e-3, 30e-3, jitter=0.1, seed=4)
    _same(gpu_ctx.mesh_tables(box.nodes_xyz, box.tets), mesh_tables(box.nodes_xyz, box.tets))


@pytest.mark.parametrize("name", ["wg_tiny", "abc_lumped", "modal_microstrip"])
def test_reference_numbering_is_kept(gpu_ctx, name):
    """edges / tris in the reference's own (CPython-set) order: every derived table equals the reference's"""
    g, t = load_golden(name)
    nodes_xyz = np.ascontiguousarray(g["nodes"].T)
    tets = np.ascontiguousarray(g["tets"].T.astype(np.int64))
    d = gpu_ctx.mesh_tables(nodes_xyz, tets, g["edges"].astype(np.int64), g["tris"].astype(np.int64))
    _same(d, t)


def test_mismatching_numbering_is_rejected(gpu_ctx):
    from emerge_b200.lib import EmergeB200Error
    g, t = load_golden("wg_tiny")
    edges = g["edges"].astype(np.int64).copy()
    edges[:, 0] = edges[:, 1]                      # duplicate edge: no longer the mesh's edge set
    with pytest.raises(EmergeB200Error):
        gpu_ctx.mesh_tables(np.ascontiguousarray(g["nodes"].T), np.ascontiguousarray(g["tets"].T.astype(np.int64)), edges,
                            g["tris"].astype(np.int64))


def test_one_million_tets_in_milliseconds(gpu_ctx):
    import time
    box = box_mesh(44, 20, 190, 22.86e-3, 10.16e-3, 98.7e-3)
    t0 = time.perf_counter()
    d = gpu_ctx.mesh_tables(box.nodes_xyz, box.tets)
    wall = time.perf_counter() - t0
    assert d.tets.shape[1] == 1003200 and d.n_field == 6484508
    # 5.6 s on the host (numpy); here sorts + lookups are milliseconds, the rest is cudaMalloc and pageable H2D of the inputs
    assert gpu_ctx.last_ms("topology") < 1500.0, gpu_ctx.last_ms("topology")
    print("topology device ms", gpu_ctx.last_ms("topology"), "wall s incl. D2H", wall)
    # spot-check against the host construction on a slice of tetrahedra (the full host build takes 6 s)
    h = mesh_tables(box.nodes_xyz, box.tets)
    _same(d, h)
