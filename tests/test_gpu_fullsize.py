"""Full-size parity properties (BASELINE config 4: 44x20x190 cells = 1,003,200 tets).  No reference result can exist at
this size (the reference's default routine raises, a direct factorisation is infeasible), so parity rests on
size-independent properties: a bitwise reproducible assembly, symmetry of the curl-curl operator, the FP64 residual of
the solve on the true A(f), and reciprocity / passivity / |S21| = 1 of the matched waveguide (analytic KAT)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_one_million_tets_properties():
    import bench
    from emerge_b200.sweep import FrequencySweep
    box, t, er, ur, bcs, L = bench.make_waveguide(44, 20, 190)
    assert t.tets.shape[1] == 1003200
    sw = FrequencySweep(t, er, ur, bcs, recycle=0)
    sw.solver_opts.update(rtol=1e-8)
    sw.f_ref = 10e9
    sw.setup()
    ctx = sw.ctx
    ns = ctx.n_solve
    rng = np.random.default_rng(0)
    x = rng.standard_normal(ns) + 1j * rng.standard_normal(ns)
    y = rng.standard_normal(ns) + 1j * rng.standard_normal(ns)
    # K alone (k0 = 0, no surface terms): complex symmetric, x^T K y = y^T K x
    ctx.form_A(0.0, [], [])
    Ky, Kx = ctx.spmv(y), ctx.spmv(x)
    a, b = np.sum(x * Ky), np.sum(y * Kx)
    assert abs(a - b) <= 1e-11 * max(abs(a), abs(b))
    # the numeric phase is bitwise reproducible: re-assemble, same operator application bit for bit
    ctx.assemble_KM()
    ctx.form_A(0.0, [], [])
    assert np.array_equal(ctx.spmv(y).view(np.float64), Ky.view(np.float64))
    # one frequency point: both ports in one block Krylov solve
    f = 10e9
    S, stats, _ = sw.solve_point(f)
    assert all(s["converged"] and s["relres"] <= 1e-8 for s in stats), stats
    assert abs(abs(S[1, 0]) - 1.0) < 0.02 and abs(abs(S[0, 1]) - 1.0) < 0.02, S
    # reciprocity holds only to the asymmetry the reference's mass matrix carries (fem/mth/tet.py:1036, SURVEY A.1: 0.4 %
    # of M): measured |S12 - S21| = 1.8e-3 here - the same defect the reference's own results have, reproduced on purpose
    assert abs(S[0, 1] - S[1, 0]) <= 1e-2, S
    assert abs(S[0, 0]) < 0.05 and abs(S[1, 1]) < 0.05, S          # matched ports
    k0 = 2 * np.pi * f / 299792458
    beta = np.sqrt(k0 ** 2 - (np.pi / bench.A_WG) ** 2)
    dphi = np.angle(S[1, 0] * np.exp(1j * beta * L), deg=True)
    assert abs(dphi) < 3.0, dphi                                   # angle(S21) = -beta L (coarse KAT, SURVEY A.13 offset)
    # the shipped tolerance is enough at full size: the same point solved cold to rtol 1e-11 gives the same S-parameters
    # within the parity target (1e-3 dB / 0.1 degrees); |S11| ~ 1e-3 is compared absolutely (util.db_deg_close)
    from tests.util import db_deg_close
    sw.solver_opts.update(rtol=1e-11)
    S_tight, stats_t, _ = sw.solve_point(f)
    assert all(s["converged"] and s["relres"] <= 1e-11 for s in stats_t), stats_t
    assert db_deg_close(S, S_tight, floor=5e-3), (S, S_tight)
    ctx.close()
    # mesh-refinement consistency: the reference's discretisation carries a density-independent phase offset (SURVEY
    # A.13); the 8x coarser mesh (125,400 tets) must give the same S21 up to discretisation error
    box2, t2, er2, ur2, bcs2, L2 = bench.make_waveguide(22, 10, 95)
    assert abs(L2 - L) < 1e-12
    sw2 = FrequencySweep(t2, er2, ur2, bcs2, recycle=0)
    sw2.f_ref = 10e9
    S_c, stats_c, _ = (sw2.setup(), sw2.solve_point(f))[1]
    assert all(s["converged"] for s in stats_c)
    assert abs(abs(S_c[1, 0]) - abs(S[1, 0])) < 3e-3, (S_c, S)
    assert abs(np.angle(S_c[1, 0] / S[1, 0], deg=True)) < 0.5, (S_c, S)
    sw2.ctx.close()
