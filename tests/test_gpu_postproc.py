"""GPU: field post-processing (csrc/postproc.cu: point location, E and curl E interpolation) against values of the
unmodified reference's EMDataSet.interpolate (tests/golden/interp_wg_tiny.npz) and, beyond the fixture, the oracle."""
import numpy as np
import pytest

from oracle import ned2_oracle as O
from tests.util import load_golden

pytestmark = pytest.mark.gpu


def _upload(ctx, t):
    ctx.upload_mesh(t.nodes, t.tets, t.tris, t.tet_to_field, t.tri_to_field, t.edges.shape[1])


def test_reference_interpolation_values(gpu_ctx):
    g, t = load_golden("interp_wg_tiny")
    _upload(gpu_ctx, t)
    pts = g["pts"]
    tet_of = gpu_ctx.locate(pts)
    assert np.array_equal(tet_of, O.locate_points(t.nodes, t.tets, pts))          # last containing tet, -1 outside
    assert np.array_equal(tet_of < 0, np.abs(g["E"]).sum(axis=0) == 0)
    const = 1.0 / (-1j * 2 * np.pi * float(g["freq"]) * (g["ur00"] * 4 * np.pi * 1e-7))      # emdata.py:193
    for tid in (None, tet_of):                   # located on the device / given by the caller
        E, H = gpu_ctx.interp_fields(g["x"], pts, tet_ids=tid, curl_const=const)
        assert np.abs(E - g["E"]).max() <= 1e-10 * np.abs(g["E"]).max()
        assert np.abs(H - g["H"]).max() <= 1e-10 * np.abs(g["H"]).max()
    E_only, none = gpu_ctx.interp_fields(g["x"], pts)
    assert none is None and np.array_equal(E_only, E)


def test_larger_mesh_against_oracle(gpu_ctx):
    from emerge_b200.synthmesh import box_mesh, mesh_tables
    box = box_mesh(8, 6, 16, 22.86e-3, 10.16e-3, 30e-3, jitter=0.15, seed=9)
    t = mesh_tables(box.nodes_xyz, box.tets)
    _upload(gpu_ctx, t)
    rng = np.random.default_rng(3)
    x = rng.standard_normal(t.n_field) + 1j * rng.standard_normal(t.n_field)
    a, b, L = box.dims
    pts = np.stack([(rng.random(3000) - 0.5) * a * 1.05, (rng.random(3000) - 0.5) * b * 1.05, (rng.random(3000) * 1.05 - 0.025) * L])
    pts[:, :200] = t.nodes[:, rng.integers(0, t.nodes.shape[1], 200)]                 # mesh vertices: many tets contain them
    cc = rng.standard_normal(t.tets.shape[1]) + 1j * rng.standard_normal(t.tets.shape[1])
    tet_ref = O.locate_points(t.nodes, t.tets, pts)
    assert np.array_equal(gpu_ctx.locate(pts), tet_ref)
    E, H = gpu_ctx.interp_fields(x, pts, curl_const=cc)
    Er, Hr = O.interp_fields(t.nodes, t.tets, t.edges, t.tris, t.tet_to_field, x, pts, tet_ref, cc)
    assert np.abs(E - Er).max() <= 1e-11 * np.abs(Er).max()
    assert np.abs(H - Hr).max() <= 1e-11 * np.abs(Hr).max()


def test_one_million_tets_20k_points(gpu_ctx):
    import bench
    box, t, er, ur, bcs, L = bench.make_waveguide(44, 20, 190)
    _upload(gpu_ctx, t)
    rng = np.random.default_rng(0)
    n = 20000
    pts = np.stack([(rng.random(n) - 0.5) * bench.A_WG, (rng.random(n) - 0.5) * bench.B_WG, rng.random(n) * L])
    tet_of = gpu_ctx.locate(pts)
    print("locate 20k points in 1M tets: ms", gpu_ctx.last_ms("locate"))
    assert (tet_of >= 0).all()
    # every located tet does contain its point (barycentric coordinates within the reference's tolerances)
    v = t.nodes[:, t.tets[:, tet_of]]                                     # (3, 4, n)
    Bm = np.stack([v[:, k] - v[:, 0] for k in (1, 2, 3)], axis=1).transpose(2, 0, 1)      # (n, 3, 3) columns
    loc = np.linalg.solve(Bm, (pts - v[:, 0]).T[:, :, None])[:, :, 0]
    assert (loc >= -1e-6).all() and (loc.sum(axis=1) <= 1.00000001).all()


def test_far_field_matches_reference_and_oracle(gpu_ctx):
    """emb_stratton_chu against fem.physics.edm.stratton_chu (tests/golden/farfield_patch.npz) and, on a larger random
    surface sample set with a field-strength cut-off that actually drops samples, against the oracle."""
    from tests.test_oracle_farfield_cpu import _g, surface_of, FF_TOL
    from emerge_b200.farfield import stratton_chu
    g = _g()
    E, H = stratton_chu(g["Ein"], g["Hin"], surface_of(g), g["theta"], g["phi"], float(g["k0"]), ctx=gpu_ctx)
    assert np.abs(E - g["E"]).max() <= FF_TOL * np.abs(g["E"]).max()
    assert np.abs(H - g["H"]).max() <= FF_TOL * np.abs(g["H"]).max()
    rng = np.random.default_rng(11)
    n, nout = 20011, 333                                   # three source chunks
    Ein = (rng.standard_normal((3, n)) + 1j * rng.standard_normal((3, n))) * rng.random(n) ** 6     # many weak samples
    Hin = (rng.standard_normal((3, n)) + 1j * rng.standard_normal((3, n))) / 377.0
    pos = rng.standard_normal((3, n)) * 0.05
    wns = rng.standard_normal((3, n)) * 1e-5
    th, ph = rng.uniform(-np.pi, np.pi, nout), rng.uniform(0, 2 * np.pi, nout)
    E, H = gpu_ctx.stratton_chu(Ein, Hin, pos, wns, th, ph, 52.3)
    Er, Hr = O.stratton_chu_ff(Ein, Hin, pos, wns, th, ph, 52.3)
    assert np.abs(E - Er).max() <= 1e-6 * np.abs(Er).max()          # both sum in FP64 on the same float32-rounded inputs
    assert np.abs(H - Hr).max() <= 1e-6 * np.abs(Hr).max()
    E2, H2 = gpu_ctx.stratton_chu(Ein, Hin, pos, wns, th, ph, 52.3)
    assert np.array_equal(E, E2) and np.array_equal(H, H2)          # fixed-order reduction: reproducible
