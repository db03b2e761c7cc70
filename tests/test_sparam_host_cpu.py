"""Host logic of the S-parameter extraction (emfreq3d.py:734-779 restated in emerge_b200/sweep.py): the batched path - one
device call per solved port for the field samples of all ports, mode fields evaluated once per frequency point - gives exactly
what the per-port-pair path gives.  The device is replaced by a deterministic stand-in for the point evaluation."""
import numpy as np
import pytest

from tests.util import golden_bcs, load_golden


class _FakeCtx:
    """E(tet, xyz) = a fixed smooth complex field + a term that depends on the tet id; counts the calls"""

    def __init__(self):
        self.calls = 0

    def interp(self, x_full, tet_ids, xyz):
        self.calls += 1
        tet_ids = np.asarray(tet_ids)
        x, y, z = np.asarray(xyz)
        E = np.empty((3, len(tet_ids)), dtype=np.complex128)
        E[0] = np.cos(40 * x) + 1j * np.sin(25 * y) + 1e-3 * tet_ids
        E[1] = np.exp(1j * 60 * z) * (1 + 30 * x) - 2e-3j * tet_ids
        E[2] = 0.3 * x * y + 1j * z
        return E


@pytest.mark.parametrize("fixture", ["wg_tiny", "abc_lumped", "modal_microstrip"])
def test_batched_sparam_samples_equal_per_port_path(fixture):
    from emerge_b200.sweep import FrequencySweep, _is
    g, t = load_golden(fixture)
    bcs = golden_bcs(g, t)
    ctx = _FakeCtx()
    sw = FrequencySweep(t, g["er"], g["ur"], bcs, ctx=ctx)
    sw.robin = [b for b in bcs if _is(b, "RobinBC")]
    sw.ports = [b for b in bcs if _is(b, "PortBC")]
    assert sw.ports
    sw._setup_sample_points()
    k0 = 2 * np.pi * 9.5e9 / 299792458
    for pa in sw.ports:
        pa.active = False
    for pa in sw.ports:                      # excited port
        pa.active = True
        ref = [sw._s_data(pb, k0, None) for pb in sw.ports]
        n0 = ctx.calls
        E_all, modes = sw._interp_all(None), {}
        assert ctx.calls == n0 + 1          # ONE device call for all ports
        got = [sw._s_data(pb, k0, None, E_all, modes) for pb in sw.ports]
        got2 = [sw._s_data(pb, k0, None, E_all, modes) for pb in sw.ports]      # second use of the cached mode fields
        assert ctx.calls == n0 + 1
        pa.active = False
        for r, a, b in zip(ref, got, got2):
            assert np.array_equal(np.asarray(r), np.asarray(a)) and np.array_equal(np.asarray(r), np.asarray(b))
