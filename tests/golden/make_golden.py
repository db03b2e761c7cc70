"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) through the
stub harness (oracle/refharness).  Run in the build container only:

    NUMBA_CACHE_DIR=/tmp/numba_cache python tests/golden/make_golden.py [case ...]

The reference has no tests or golden vectors of its own, so every pin is an output of
the reference itself on a deterministic synthetic mesh.  Each fixture stores the mesh in the
reference's own numbering (nodes/tets/edges/tris; all other tables are derived from these and
asserted equal to the reference's), the inputs (er, ur, BC description) and reference outputs.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
from oracle.refharness import harness as H  # noqa: E402

H.setup_paths()
from emerge_b200.synthmesh import box_mesh, mesh_tables  # noqa: E402

WR90 = (22.86e-3, 10.16e-3)


def _tables_check(mesh, basis):
    t = mesh_tables(np.ascontiguousarray(mesh.nodes.T), np.ascontiguousarray(mesh.tets.T), mesh.edges, mesh.tris)
    for name in ("tet_to_edge", "tet_to_tri", "tri_to_edge", "tri_to_tet"):
        assert np.array_equal(getattr(t, name), getattr(mesh, name)), name
    assert np.allclose(t.edge_lengths, mesh.edge_lengths, rtol=0, atol=0)
    assert np.array_equal(t.tet_to_field, basis.tet_to_field)
    assert np.array_equal(t.tri_to_field, basis.tri_to_field)
    assert np.array_equal(t.edge_to_field, basis.edge_to_field)
    return t


def _mesh_dict(mesh):
    return dict(nodes=mesh.nodes.copy(), tets=mesh.tets.astype(np.int32), edges=mesh.edges.astype(np.int32),
                tris=mesh.tris.astype(np.int32))


def _csr_dict(prefix, M):
    M = M.tocsr()
    M.sort_indices()
    return {prefix + "_indptr": M.indptr.astype(np.int64), prefix + "_indices": M.indices.astype(np.int32),
            prefix + "_data": M.data.copy()}


def _sweep(fem, phys, mesh, freqs, full, out, keep_x=()):
    """Run frequency_domain() and collect outputs; `full` stores matrices/fields, else only checks."""
    from fem.elements.nedelec2 import Nedelec2
    phys.frequencies = list(freqs)
    data = phys.frequency_domain()
    basis = phys.basis
    _tables_check(mesh, basis)
    E, B = phys.assembler.cached_matrices
    er = mesh.retreive(lambda mat, x, y, z: mat.fer3d_mat(x, y, z), phys.mesher.volumes)
    ur = mesh.retreive(lambda mat, x, y, z: mat.fur3d_mat(x, y, z), phys.mesher.volumes)
    out.update(_mesh_dict(mesh))
    out["er"], out["ur"] = er, ur
    out["freqs"] = np.asarray(freqs, dtype=np.float64)
    rng = np.random.default_rng(1234)
    v = rng.standard_normal(E.shape[0]) + 1j * rng.standard_normal(E.shape[0])
    out["probe_v"] = v
    out["E_dot_v"], out["B_dot_v"] = E @ v, B @ v
    out["E_nnz"] = np.int64(E.nnz)
    out["E_absmax"], out["B_absmax"] = np.abs(E.data).max(), np.abs(B.data).max()
    if full:
        out.update(_csr_dict("E", E))
        out["B_data"] = B.tocsr().data.copy()
        assert np.array_equal(E.indptr, B.indptr) and np.array_equal(E.indices, B.indices)
    ports = [bc for bc in phys.boundary_conditions if isinstance(bc, fem.bc.PortBC)]
    out["port_numbers"] = np.array([p.port_number for p in ports], dtype=np.int64)
    S = np.zeros((len(freqs), len(ports), len(ports)), dtype=np.complex128)
    for i, f in enumerate(freqs):
        ds = data.item(i)
        S[i] = ds.Sp.arry
        K, b, solve_ids, pv = phys.assembler.assemble_freq_matrix(basis, er, ur, phys.boundary_conditions, f,
                                                                   cache_matrices=True)
        if i == 0:
            out["solve_ids"] = np.asarray(solve_ids, dtype=np.int64)
        out[f"K_dot_v_{i}"] = K @ v
        for p in ports:
            out[f"bvec_{i}_p{p.port_number}"] = pv[p.port_number]
            x = ds._fields[p.port_number]
            r = K[np.ix_(solve_ids, solve_ids)] @ x[solve_ids] - pv[p.port_number][solve_ids]
            out[f"xres_{i}_p{p.port_number}"] = np.linalg.norm(r) / np.linalg.norm(pv[p.port_number][solve_ids])
            if full or i in keep_x:
                out[f"x_{i}_p{p.port_number}"] = x
        if full and i == 0:
            K.eliminate_zeros()
            out.update(_csr_dict("K0", K))
    out["S"] = S
    return data


def case_wg_tiny():
    """3x2x4 cells (144 tets), jittered, vacuum, two RectangularWaveguide ports + PEC walls; everything stored."""
    box = box_mesh(3, 2, 4, *WR90, 20e-3, jitter=0.1, seed=0)
    fem, phys, mesh = H.build_physics(box)
    H.rect_waveguide_ports(fem, phys, box)
    out = dict(kind="rectwg", dims=np.array(box.dims), face_tris=box.face_tris.astype(np.int32), face_tag=box.face_tag)
    _sweep(fem, phys, mesh, [9e9, 10e9], True, out)
    # element-level pins: the reference's element matrices for the first 6 tets
    from fem.mth.tet import ned2_tet_stiff_mass
    from fem.mth.optimized import matinv
    from fem.physics.edm.optimized_assembly import local_tet_to_triid, local_tet_to_edgeid
    basis = phys.basis
    Es, Bs = [], []
    for it in range(6):
        ltm = local_tet_to_triid(basis.tet_to_field, mesh.tets, mesh.tris, it, mesh.n_edges)
        lem = local_tet_to_edgeid(mesh.tets, mesh.edges, basis.tet_to_field, it)
        Esub, Bsub = ned2_tet_stiff_mass(mesh.nodes[:, mesh.tets[:, it]], mesh.edge_lengths[mesh.tet_to_edge[:, it]],
                                         lem, ltm, matinv(out["ur"][:, :, it]), out["er"][:, :, it])
        Es.append(Esub), Bs.append(Bsub)
    out["elemE"], out["elemB"] = np.array(Es), np.array(Bs)
    # full (non-diagonal, non-symmetric) tensors on the same tets: pins matinv's adj*det quirk (App. A.2)
    rng = np.random.default_rng(7)
    Es, Bs, ers, urs = [], [], [], []
    for it in range(4):
        ert = rng.standard_normal((3, 3)) + 1j * rng.standard_normal((3, 3)) + 3 * np.eye(3)
        urt = rng.standard_normal((3, 3)) + 1j * rng.standard_normal((3, 3)) + 3 * np.eye(3)
        ltm = local_tet_to_triid(basis.tet_to_field, mesh.tets, mesh.tris, it, mesh.n_edges)
        lem = local_tet_to_edgeid(mesh.tets, mesh.edges, basis.tet_to_field, it)
        Esub, Bsub = ned2_tet_stiff_mass(mesh.nodes[:, mesh.tets[:, it]], mesh.edge_lengths[mesh.tet_to_edge[:, it]],
                                         lem, ltm, matinv(urt), ert)
        Es.append(Esub), Bs.append(Bsub), ers.append(ert), urs.append(urt)
    out["full_er"], out["full_ur"] = np.array(ers), np.array(urs)
    out["full_elemE"], out["full_elemB"] = np.array(Es), np.array(Bs)
    return out


def _mat_fn(diag):
    d = np.asarray(diag, dtype=np.complex128)

    def fn(x, y, z):
        return np.repeat(np.diag(d)[:, :, None], x.shape[0], axis=2)
    return fn


def case_wg_materials():
    """3x3x4 cells, two volumes: lossy isotropic dielectric and a diagonal-anisotropic, magnetic medium."""
    a, b = WR90
    L = 18e-3
    box = box_mesh(3, 3, 4, a, b, L, jitter=0.12, seed=3, vol_fn=lambda x, y, z: np.where(z < L / 2, 1, 2))
    H.setup_paths()
    import fem
    m1 = fem.Material(er=2.2, tand=0.01)
    m2 = fem.Material(_fer=_mat_fn([3.0 - 0.2j, 1.5, 2.0 - 0.05j]), _fur=_mat_fn([1.2, 0.9 - 0.1j, 1.0]))
    fem, phys, mesh = H.build_physics(box, {1: m1, 2: m2})
    H.rect_waveguide_ports(fem, phys, box)
    out = dict(kind="rectwg", dims=np.array(box.dims), face_tris=box.face_tris.astype(np.int32), face_tag=box.face_tag)
    _sweep(fem, phys, mesh, [8.5e9, 11e9], True, out)
    return out


def case_wg_medium():
    """8x4x16 cells (3072 tets, N~20k), vacuum waveguide: S-parameters + matrix checksums only."""
    box = box_mesh(8, 4, 16, *WR90, 40e-3, jitter=0.08, seed=1)
    fem, phys, mesh = H.build_physics(box)
    H.rect_waveguide_ports(fem, phys, box)
    out = dict(kind="rectwg", dims=np.array(box.dims), face_tris=box.face_tris.astype(np.int32), face_tag=box.face_tag)
    _sweep(fem, phys, mesh, [8e9, 9e9, 10e9, 11e9, 12e9], False, out)
    for k in [k for k in out if k.startswith(("K_dot_v_", "bvec_")) and not k.startswith(("K_dot_v_0", "bvec_0"))]:
        del out[k]
    return out


def abc_lumped_physics():
    """Patch-like box: dielectric slab, internal PEC patch, vertical LumpedPort plate, AbsorbingBoundary on
    5 outer faces, PEC ground (look-alike of demo3_patch_antenna.py, SURVEY App. C.7 iii)."""
    a, b, L = 24e-3, 24e-3, 12e-3           # x,y footprint; z height
    nx, ny, nz = 6, 6, 4
    hz = L / nz
    hx = a / nx

    def vol(x, y, z):
        return np.where(z < hz, 1, 2)        # slab of one cell height

    def patch(x, y, z):                       # PEC patch on the slab top, central 2x2 cells
        return (np.abs(z - hz) < 1e-9) & (np.abs(x) < hx) & (np.abs(y) < hx)

    def plate(x, y, z):                       # vertical plate in plane y=0... use plane x = -hx, z<hz, |y|<hx
        return (np.abs(x + hx) < 1e-9) & (z < hz) & (np.abs(y) < hx)

    box = box_mesh(nx, ny, nz, a, b, L, jitter=0.0, seed=0, vol_fn=vol,
                   internal_faces=[(7, (0, 0, 1), patch), (8, (1, 0, 0), plate)])
    H.setup_paths()
    import fem
    fem, phys, mesh = H.build_physics(box, {1: fem.Material(er=3.38, tand=0.002), 2: fem.AIR}, pec_extra_tags=(7,))
    port = fem.bc.LumpedPort(fem.FaceSelection([8]), 1, width=2 * hx, height=hz, direction=fem.ZAX, active=True, Z0=50)
    abc = fem.bc.AbsorbingBoundary(fem.FaceSelection([1, 2, 3, 4, 6]))
    phys.assign(port, abc)
    return fem, phys, mesh, box, port, hx, hz


def case_abc_lumped():
    fem, phys, mesh, box, port, hx, hz = abc_lumped_physics()
    out = dict(kind="abc_lumped", dims=np.array(box.dims), face_tris=box.face_tris.astype(np.int32),
               face_tag=box.face_tag, lumped=np.array([2 * hx, hz, 50.0]),
               vint_start=np.zeros(3), vint_end=np.zeros(3))
    _sweep(fem, phys, mesh, [2.0e9, 2.4e9], True, out)
    out["vint_start"], out["vint_end"] = [np.asarray(v, dtype=float) for v in port.voltage_integration_points]
    out["port_cs_basis"] = port.cs._basis.copy()
    out["port_cs_origin"] = np.asarray(port.cs.origin, dtype=float)
    return out


def microstrip_box(nx=12, ny=6, nz=8):
    """Shielded microstrip look-alike of demo1_stepped_imp_filter.py (SURVEY App. C.7 ii): 12 x 6 x 20 mm PEC box, lossy
    substrate (2 cells high), PEC strip 4 mm wide on the substrate top over the whole length, ports on z=0 / z=L."""
    a, b, L = 12e-3, 6e-3, 20e-3
    hy = b / ny
    ysub = -b / 2 + 2 * hy

    def vol(x, y, z):
        return np.where(y < ysub, 1, 2)

    def strip(x, y, z):
        return (np.abs(y - ysub) < 1e-9) & (np.abs(x) < 2e-3)

    return box_mesh(nx, ny, nz, a, b, L, jitter=0.0, seed=0, vol_fn=vol, internal_faces=[(7, (0, 1, 0), strip)])


def modal_physics(box, tand=0.01):
    H.setup_paths()
    import fem
    fem, phys, mesh = H.build_physics(box, {1: fem.Material(er=2.2, tand=tand), 2: fem.AIR}, pec_extra_tags=(7,))
    a, b, L = box.dims
    ports = []
    for num, (tag, z) in enumerate([(5, 0.0), (6, L)], start=1):
        cs = fem.CoordinateSystem(fem.XAX, fem.YAX, fem.ZAX, origin=np.array([0.0, 0.0, z]))
        ports.append(fem.bc.ModalPort(fem.FaceSelection([tag]), num, cs=cs))
    phys.assign(*ports)
    # ParallelRoutine (the reference's SuperLU routine) carries no eigen-solver; the reference's default routine uses
    # SolverLAPACK for modal_analysis(direct=True) (fem/solver.py:578-583)
    from fem.solver import SolverLAPACK
    phys.solveroutine.direct_eig_solver = SolverLAPACK()
    return fem, phys, mesh, ports


def case_modal_microstrip():
    """Two TEM ModalPorts whose mode comes from the reference's own modal_analysis (emfreq3d.py:201-364,
    bc.py:329-494): BASELINE configs 1-2 (demo1 / demo2 use ModalPort).  The mode field is stored sampled at the
    Dunavant-4 points of the port triangles (the only points the hot path evaluates it at)."""
    from emerge_b200.sweep import dunavant4
    box = microstrip_box()
    fem, phys, mesh, ports = modal_physics(box)
    freqs = [1e9, 2e9, 3e9]
    phys.frequencies = list(freqs)
    for p in ports:
        phys.modal_analysis(p, 1, direct=True, TEM=True, freq=freqs[0])
    out = dict(kind="modal_microstrip", dims=np.array(box.dims), face_tris=box.face_tris.astype(np.int32),
               face_tag=box.face_tag)
    _sweep(fem, phys, mesh, freqs, False, out, keep_x=(0,))     # matrices as checksums; fields of the first point
    for k in [k for k in out if k.startswith("K_dot_v_") and k != "K_dot_v_0"]:
        del out[k]
    DP = dunavant4()
    for p in ports:
        ids = mesh.get_triangles(p.tags)
        P = mesh.nodes[:, mesh.tris[:, ids]]                        # (3 xyz, 3 vert, ntri)
        pts = np.einsum("kq,xkt->xqt", DP[1:4], P).reshape(3, -1)
        n = p.port_number
        mode = p.get_mode()
        out[f"mode_pts_p{n}"] = pts
        out[f"mode_E_p{n}"] = np.asarray(p.port_mode_3d_global(pts[0], pts[1], pts[2], 1.0), dtype=np.complex128)
        out[f"mode_scalars_p{n}"] = np.array([mode.beta, mode.k0, float(bool(mode.TEM)), mode.freq, mode.norm_factor,
                                              complex(mode.Z0).real, complex(mode.Z0).imag], dtype=np.float64)
        out[f"mode_type_p{n}"] = str(mode.modetype)
    return out


def case_lossy_slabs():
    """Dielectric-loaded waveguide (look-alike of BASELINE config 5 at fixture size): eps_r = 9.8(1 - 1e-4j) slabs periodic
    in z, vacuum elsewhere, RectangularWaveguide ports.  Pins the solver on a complex (lossy) mass matrix."""
    a, b = WR90
    L = 36e-3
    nz = 18
    hz = L / nz

    def vol(x, y, z):
        return np.where((np.floor(z / hz).astype(int) % 6) >= 4, 2, 1)      # 2 of every 6 cell layers are ceramic

    box = box_mesh(6, 3, nz, a, b, L, jitter=0.08, seed=5, vol_fn=vol)
    H.setup_paths()
    import fem
    cer = fem.Material(_fer=_mat_fn([9.8 * (1 - 1e-4j)] * 3))
    fem, phys, mesh = H.build_physics(box, {1: fem.VACUUM, 2: cer})
    H.rect_waveguide_ports(fem, phys, box)
    out = dict(kind="rectwg", dims=np.array(box.dims), face_tris=box.face_tris.astype(np.int32), face_tag=box.face_tag)
    _sweep(fem, phys, mesh, [8e9, 9.3e9, 10.1e9, 12e9], False, out)
    for k in [k for k in out if k.startswith(("K_dot_v_", "bvec_")) and not k.startswith(("K_dot_v_0", "bvec_0"))]:
        del out[k]
    return out


def case_lossy_slabs_50k():
    """Config-5 look-alike at the largest size the reference's direct solver handles in minutes here (49,680 tets, 320k
    dofs): ceramic slabs eps_r = 9.8 (1 - 1e-4j) in two of every twelve cell layers (port planes in vacuum), three
    frequencies.  Slim fixture: mesh in the reference's numbering, the slab flag per tet, S-parameters and the reference's
    own residuals - the pin for the iterative solver on a lossy, strongly resonant operator beyond toy size."""
    a, b = WR90
    nx, ny, nz = 20, 9, 46
    L = nz * a / nx
    hz = L / nz

    def vol(x, y, z):
        layer = np.floor(z / hz).astype(np.int64) % 12
        return np.where((layer == 5) | (layer == 6), 2, 1)

    box = box_mesh(nx, ny, nz, a, b, L, jitter=0.05, seed=11, vol_fn=vol)
    H.setup_paths()
    import fem
    cer = fem.Material(_fer=_mat_fn([9.8 * (1 - 1e-4j)] * 3))
    fem, phys, mesh = H.build_physics(box, {1: fem.VACUUM, 2: cer})
    H.rect_waveguide_ports(fem, phys, box)
    freqs = [8.5e9, 10e9, 11.5e9]
    full = {}
    _sweep(fem, phys, mesh, freqs, False, full)
    out = dict(kind="rectwg", dims=np.array(box.dims), face_tris=box.face_tris.astype(np.int32), face_tag=box.face_tag)
    out.update(_mesh_dict(mesh))
    ceramic = np.abs(full["er"][0, 0] - 1.0) > 1e-12
    out.update(ceramic=ceramic, eps_ceramic=np.complex128(9.8 * (1 - 1e-4j)), freqs=full["freqs"], S=full["S"],
               port_numbers=full["port_numbers"], solve_ids_len=np.int64(len(full["solve_ids"])),
               xres=np.array([[full[f"xres_{i}_p{p}"] for p in full["port_numbers"]] for i in range(len(freqs))]))
    return out


def case_interp_wg_tiny():
    """Field post-processing pins: EMDataSet.interpolate (emdata.py:181-199) -> ned2_tet_interp /
    ned2_tet_interp_curl (mth/tet.py:371-626) on the wg_tiny mesh at random points (some outside the mesh): E and H of the
    port-1 solution.  Stores the field vector too, so the test needs no solve."""
    box = box_mesh(3, 2, 4, *WR90, 20e-3, jitter=0.1, seed=0)
    fem, phys, mesh = H.build_physics(box)
    H.rect_waveguide_ports(fem, phys, box)
    phys.frequencies = [9e9]
    data = phys.frequency_domain()
    ds = data.item(0)
    rng = np.random.default_rng(42)
    a, b, L = box.dims
    n = 160
    pts = np.stack([(rng.random(n) - 0.5) * a * 1.15, (rng.random(n) - 0.5) * b * 1.15, (rng.random(n) * 1.15 - 0.075) * L])
    # a few points exactly on vertices / faces of the mesh: the "last containing tetrahedron wins" rule decides
    pts[:, :6] = mesh.nodes[:, [5, 17, 23, 31, 40, 47]]
    pts[:, 6:10] = mesh.nodes[:, mesh.tris[:, [3, 30, 60, 90]]].mean(axis=1)
    ds.interpolate(pts[0], pts[1], pts[2])
    out = dict(_mesh_dict(mesh))
    out.update(kind="rectwg", dims=np.array(box.dims), face_tris=box.face_tris.astype(np.int32), face_tag=box.face_tag,
               freq=np.float64(9e9), x=np.asarray(ds._field), ur00=np.asarray(ds.ur), pts=pts,
               E=np.array([ds.Ex, ds.Ey, ds.Ez]), H=np.array([ds.Hx, ds.Hy, ds.Hz]))
    return out


def case_bma_microstrip():
    """Boundary-mode analysis pins: the reference's Assembler.assemble_bma_matrices (assembler.py:246-308)
    on port 1 of the microstrip look-alike - inputs (SurfaceMesh tables, port-local coordinates, er/ur per port triangle),
    the assembled E, B (CSR), solve_ids - plus element matrices straight from generalized_matrix_GQ for a few triangles
    with FULL non-symmetric complex tensors, and the eigenpair modal_analysis(direct=True, TEM=True) selects."""
    from fem.physics.edm.nedeleclegrange2 import generalized_matrix_GQ, local_tri_to_edgeid
    from fem.mth.optimized import matinv
    box = microstrip_box()
    fem, phys, mesh, ports = modal_physics(box)
    freq = 1e9
    phys.frequencies = [freq]
    port = ports[0]
    phys.modal_analysis(port, 1, direct=True, TEM=True, freq=freq)
    k0 = 2 * np.pi * freq / 299792458
    er, ur = port._er, port._ur                                   # (3,3,n_tris of the whole mesh), emfreq3d.py:322-323
    E, B, solve_ids, nlf = phys.assembler.assemble_bma_matrices(phys.basis, er, ur, k0, port, phys.boundary_conditions)
    sm = nlf.mesh
    tri_ids = mesh.get_triangles(port.tags)
    xy = (np.linalg.pinv(port.cs._basis) @ sm.nodes)[:2]
    out = dict(kind="bma", k0=np.float64(k0), xy=xy, s_tris=sm.tris.astype(np.int64), s_edges=sm.edges.astype(np.int64),
               s_tri_to_edge=sm.tri_to_edge.astype(np.int64), er=er[:, :, tri_ids], ur=ur[:, :, tri_ids],
               n_nodes=np.int64(sm.n_nodes), tri_to_field=nlf.tri_to_field.astype(np.int64), n_field=np.int64(nlf.n_field),
               solve_ids=np.asarray(solve_ids, dtype=np.int64))
    out.update(_csr_dict("E", E))
    out.update(_csr_dict("B", B))
    mode = port.get_mode()
    out["beta"] = np.float64(mode.beta)
    out["mode_field"] = np.asarray(mode.modefield, dtype=np.complex128) if hasattr(mode, "modefield") else np.zeros(0)
    # element level, full tensors
    rng = np.random.default_rng(7)
    sel = np.array([0, 3, 11, 20])
    fer = rng.standard_normal((4, 3, 3)) + 1j * rng.standard_normal((4, 3, 3)) + 3 * np.eye(3)
    fur = rng.standard_normal((4, 3, 3)) * 0.3 + 0.2j * rng.standard_normal((4, 3, 3)) + 2 * np.eye(3)
    eA, eB = [], []
    tri_to_edge = nlf.tri_to_field[:3, :].astype(np.int64)
    for k, it in enumerate(sel):
        lm = local_tri_to_edgeid(int(it), sm.tris.astype(np.int64), sm.edges.astype(np.int64), tri_to_edge)
        a_, b_ = generalized_matrix_GQ(np.ascontiguousarray(xy[:, sm.tris[:, it]]).astype(np.float64), lm,
                                       matinv(np.ascontiguousarray(fur[k])), np.ascontiguousarray(fer[k]), float(k0))
        eA.append(a_)
        eB.append(b_)
    out.update(full_sel=sel, full_er=fer, full_ur=fur, full_A=np.array(eA), full_B=np.array(eB))
    return out


def case_farfield_patch():
    """Far-field pin: the demo3 flow (demo3_patch_antenna.py:94-100) on the abc_lumped look-alike -
    SurfaceMesh of the absorbing boundary, E/H at its edge midpoints, fem.physics.edm.stratton_chu over a theta cut and a
    phi cut.  Stores the SurfaceMesh arrays the function reads, so the test needs neither mesh nor solve."""
    fem, phys, mesh, box, port, hx, hz = abc_lumped_physics()
    phys.frequencies = [2.4e9]
    data = phys.frequency_domain()
    surf = mesh.boundary_surface([1, 2, 3, 4, 6], (0, 0, box.dims[2] / 2))
    ds = data.item(0)
    Ein, Hin = ds.interpolate(*surf.exyz).EH
    theta = np.concatenate([np.linspace(-np.pi, np.pi, 61), np.full(24, 0.3)])
    phi = np.concatenate([np.zeros(61), np.linspace(0, 2 * np.pi, 24)])
    from fem.physics.edm import stratton_chu
    E, Hf = stratton_chu(Ein, Hin, surf, theta, phi, ds.k0)
    return dict(kind="farfield", Ein=np.asarray(Ein), Hin=np.asarray(Hin), areas=np.asarray(surf.areas),
                normals=np.asarray(surf.normals), tri_to_edge=np.asarray(surf.tri_to_edge).astype(np.int32),
                edge_centers=np.asarray(surf.edge_centers), theta=theta, phi=phi, k0=np.float64(ds.k0), E=E, H=Hf)


CASES = dict(wg_tiny=case_wg_tiny, wg_materials=case_wg_materials, wg_medium=case_wg_medium,
             abc_lumped=case_abc_lumped, modal_microstrip=case_modal_microstrip, lossy_slabs=case_lossy_slabs,
             interp_wg_tiny=case_interp_wg_tiny, farfield_patch=case_farfield_patch, bma_microstrip=case_bma_microstrip, lossy_slabs_50k=case_lossy_slabs_50k)

if __name__ == "__main__":
    if not os.path.isdir("/root/reference/fem"):
        raise SystemExit("fixtures are generated from /root/reference in the build container only")
    names = sys.argv[1:] or list(CASES)
    for n in names:
        out = CASES[n]()
        path = os.path.join(HERE, n + ".npz")
        np.savez_compressed(path, **out)
        print(n, "->", path, f"{os.path.getsize(path) / 1e6:.2f} MB", "S[0]=", out["S"][0].ravel()[:4] if "S" in out else "-")
