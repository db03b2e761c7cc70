"""Drop-in proof on the GPU box: the UNMODIFIED reference (oracle/_ref, copied by oracle/make_ref.py) builds a real
Electrodynamics3D and runs its own frequency_domain() three ways on the same object:

  1. stock CPU path: Assembler + ParallelRoutine (RCM + SuperLU), emfreq3d.py:607-732;
  2. after emerge_b200.dropin.install(physics): the reference's own loop on top of GpuAssembler.assemble_freq_matrix and
     the patched SolveRoutine.solve (the two seams of SURVEY 8b);
  3. after install(physics, fast=True): physics.frequency_domain() / frequency_domain_par(njobs) replaced by the
     FrequencySweep-backed drivers that fill the reference's EMSimData.

EMSimData.Sp must agree within 1e-3 dB / 0.1 degrees, _fields to 1e-7, for RectangularWaveguide ports, LumpedPort +
AbsorbingBoundary, and ModalPorts whose mode comes from the reference's own modal_analysis.  The reference's result API
(axis access) and the Touchstone export (emerge_b200.touchstone; the reference's own needs scikit-rf, absent here) run on
the GPU-filled EMSimData.  Skipped when oracle/_ref is absent."""
import os

import numpy as np
import pytest

from tests.util import db_deg_close

pytestmark = pytest.mark.gpu
REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "fem")


def _builders():
    from tests.golden import make_golden as G
    from oracle.refharness import harness as H
    from emerge_b200.synthmesh import box_mesh

    def rectwg():
        box = box_mesh(5, 3, 8, *G.WR90, 24e-3, jitter=0.1, seed=2)
        fem, phys, mesh = H.build_physics(box)
        H.rect_waveguide_ports(fem, phys, box)
        return phys, [8.5e9, 10e9, 11.5e9]

    def abc_lumped():
        fem, phys, mesh, box, port, hx, hz = G.abc_lumped_physics()
        return phys, [2.0e9, 2.2e9, 2.4e9]

    def modal():
        box = G.microstrip_box(8, 6, 5)
        fem, phys, mesh, ports = G.modal_physics(box)
        phys.frequencies = [1e9]
        for p in ports:
            phys.modal_analysis(p, 1, direct=True, TEM=True, freq=1e9)
        return phys, [1e9, 2e9, 3e9]
    return dict(rectwg=rectwg, abc_lumped=abc_lumped, modal=modal)


def _interp(data, i, seed=0):
    """EMDataSet.interpolate at 200 random points around the mesh (emdata.py:181-199): (E (3,n), H (3,n))"""
    ds = data.item(i)
    nodes = ds._basis.mesh.nodes
    lo, hi = nodes.min(axis=1), nodes.max(axis=1)
    rng = np.random.default_rng(seed)
    pts = lo[:, None] + (hi - lo)[:, None] * (rng.random((3, 200)) * 1.1 - 0.05)
    ds.interpolate(pts[0], pts[1], pts[2])
    return np.array([ds.Ex, ds.Ey, ds.Ez]), np.array([ds.Hx, ds.Hy, ds.Hz])


def _far_field(phys, data, reference=False):
    """demo3_patch_antenna.py:94-100: far field of the absorbing boundary's surface mesh over a theta cut.
    reference: the reference's own function even when an earlier install() in this process has replaced it"""
    import fem.physics.edm as edm
    fn = getattr(edm.stratton_chu, "_reference", edm.stratton_chu) if reference else edm.stratton_chu
    surf = phys.mesh.boundary_surface([1, 2, 3, 4, 6], (0.0, 0.0, 6e-3))
    ds = data.item(0)
    Ein, Hin = ds.interpolate(*surf.exyz).EH
    theta = np.linspace(-np.pi, np.pi, 73)
    return fn(Ein, Hin, surf, theta, 0 * theta + 0.2, ds.k0)


def _check_modal_analysis(phys, asm):
    """Electrodynamics3D.modal_analysis (emfreq3d.py:201-364) with GpuAssembler.assemble_bma_matrices underneath: the
    matrices equal the stock assembler's to round-off and the re-computed mode has the same propagation constant,
    impedance and (sampled) field as the mode the stock path found."""
    import fem.physics.edm.assembler as ref_asm
    from fem.bc import ModalPort
    ports = [b for b in phys.boundary_conditions if isinstance(b, ModalPort)]
    port = ports[0]
    old = port.get_mode()
    k0 = 2 * np.pi * 1e9 / 299792458
    E0, B0, ids0, _ = ref_asm.Assembler().assemble_bma_matrices(phys.basis, port._er, port._ur, k0, port, phys.boundary_conditions)
    E1, B1, ids1, nlf = asm.assemble_bma_matrices(phys.basis, port._er, port._ur, k0, port, phys.boundary_conditions)
    assert np.array_equal(ids0, ids1)
    assert abs(E1 - E0).max() <= 1e-12 * abs(E0).max() and abs(B1 - B0).max() <= 1e-12 * abs(B0).max()
    tri = phys.mesh.get_triangles(port.tags)
    c = phys.mesh.nodes[:, phys.mesh.tris[:, tri]].mean(axis=1)
    F_old = np.asarray(port.port_mode_3d_global(c[0], c[1], c[2], 1.0))
    port.modes = []
    phys.modal_analysis(port, 1, direct=True, TEM=True, freq=1e9)           # the reference driver on the device element loop
    new = port.get_mode()
    assert abs(new.beta - old.beta) <= 1e-9 * abs(old.beta)
    assert abs(complex(new.Z0) - complex(old.Z0)) <= 1e-7 * abs(complex(old.Z0))
    F_new = np.asarray(port.port_mode_3d_global(c[0], c[1], c[2], 1.0))
    assert np.abs(np.abs(F_new) - np.abs(F_old)).max() <= 1e-7 * np.abs(F_old).max()


def _collect(data, nf):
    S = np.array([data.item(i).Sp.arry.copy() for i in range(nf)])
    fields = [{k: np.array(v) for k, v in data.item(i)._fields.items()} for i in range(nf)]
    return S, fields


def _check(data, nf, S_ref, F_ref, what, fields=True):
    S, F = _collect(data, nf)
    assert db_deg_close(S, S_ref), (what, S, S_ref)
    if fields:
        for i in range(nf):
            assert set(F[i]) == set(F_ref[i]), what
            for k, x in F[i].items():
                assert np.linalg.norm(x - F_ref[i][k]) <= 1e-7 * np.linalg.norm(F_ref[i][k]), (what, i, k)
    for i in range(nf):
        d, r = data.item(i), None
        assert d.er.shape == d.ur.shape and len(d.port_modes) == S.shape[1]


@pytest.mark.skipif(not os.path.isdir(REF), reason="oracle/_ref (copy of the reference package) not present")
@pytest.mark.parametrize("case", ["rectwg", "abc_lumped", "modal"])
def test_reference_frequency_domain_on_top_of_install(case, tmp_path):
    from emerge_b200.dropin import install, GpuAssembler
    phys, freqs = _builders()[case]()
    nf = len(freqs)
    phys.frequencies = list(freqs)
    data = phys.frequency_domain()                                   # 1. the reference, untouched
    S_ref, F_ref = _collect(data, nf)
    E_ref, H_ref = _interp(data, 0)                                  # reference post-processing (numba, all tets x all points)
    stock_solve = type(phys.solveroutine).solve
    ff_ref = _far_field(phys, data, reference=True) if case == "abc_lumped" else None      # reference Stratton-Chu (sc.py), demo3 flow

    asm = install(phys, rtol=1e-10)                                  # 2. the two seams
    assert isinstance(phys.assembler, GpuAssembler)
    if case == "modal":                                              # boundary-mode analysis on top of the device element loop
        _check_modal_analysis(phys, asm)
    data = phys.frequency_domain()
    _check(data, nf, S_ref, F_ref, "seams")
    data = phys.frequency_domain()                                   # same problem again: the device state is reused
    _check(data, nf, S_ref, F_ref, "seams, second run")
    asm.ctx.close()

    phys.solveroutine.solve = stock_solve.__get__(phys.solveroutine)  # 3. the fast drivers, shipped tolerance
    asm = install(phys, fast=True)
    assert asm.solver_opts["rtol"] == 1e-8
    data = phys.frequency_domain()
    _check(data, nf, S_ref, F_ref, "fast driver")
    assert getattr(phys.basis, "_emb_postproc", False)
    E, Hf = _interp(data, 0)                                         # same call, now located + evaluated on the device
    assert np.array_equal(np.abs(E).sum(axis=0) == 0, np.abs(E_ref).sum(axis=0) == 0)       # same points outside the mesh
    assert np.abs(E - E_ref).max() <= 1e-6 * np.abs(E_ref).max() and np.abs(Hf - H_ref).max() <= 1e-6 * np.abs(H_ref).max()
    if ff_ref is not None:                                           # fem.physics.edm.stratton_chu now runs emb_stratton_chu
        import fem.physics.edm as edm
        assert edm.stratton_chu.__module__ == "emerge_b200.farfield"
        ff = _far_field(phys, data)
        for a, b in zip(ff, ff_ref):
            assert np.abs(a - b).max() <= 2e-5 * np.abs(b).max()
    data = phys.frequency_domain_par(njobs=2)
    _check(data, nf, S_ref, F_ref, "fast parallel driver (one rank)")
    # result API of the reference on the GPU-filled object: axis access and Touchstone export (emdata.py:284-331)
    f_ax, s21 = data.ax("freq").S(1, 1)
    assert np.allclose(f_ax, freqs) and np.allclose(s21, S_ref[:, 0, 0], atol=1e-4)
    from emerge_b200.touchstone import export_touchstone, read_touchstone
    f_ts, S_ts, _ = read_touchstone(export_touchstone(data, str(tmp_path / "sweep"), "RI"))
    assert np.allclose(f_ts, freqs) and db_deg_close(S_ts, S_ref)
    asm.ctx.close()
