"""GPU: the top-level auxiliary spaces G (P2 gradients) and P (Whitney) built on the device (csrc/auxbuild.cu) against the
numpy builder emerge_b200/auxspace.py::build_aux_spaces_paired (itself pinned to the generic construction in
tests/test_auxspace_cpu.py): same matrices entry for entry, same dropped columns, and the sweep on top of either builder
takes the same iterations and returns the same S-parameters."""
import numpy as np
import pytest

from tests.util import load_golden, golden_bcs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["wg_medium", "abc_lumped"])
def test_device_built_spaces_equal_the_numpy_builder(name):
    from emerge_b200.sweep import FrequencySweep
    from emerge_b200.auxspace import build_aux_spaces_paired
    g, t = load_golden(name)
    sw = FrequencySweep(t, g["er"], g["ur"], golden_bcs(g, t), recycle=0)
    sw.f_ref = float(np.median(g["freqs"]))
    sw.setup()                                              # device builder (default): spaces 0 (G) and 1 (P)
    ctx = sw.ctx
    assert ctx.paired
    keep = np.ones(t.n_field, dtype=bool)
    keep[sw.pec_ids] = False
    Gs, Ps, badP, _ = build_aux_spaces_paired(t, keep)
    for idx, ref in ((0, Gs), (1, Ps)):
        M = ctx.aux_get(idx)
        ref = ref.tocsr()
        ref.sort_indices()
        assert M.shape == ref.shape and np.array_equal(M.indptr, ref.indptr) and np.array_equal(M.indices, ref.indices)
        assert np.array_equal(M.data, ref.data)             # same arithmetic: bit for bit
    ctx.close()


def test_sweep_is_unchanged_by_the_builder(monkeypatch):
    from emerge_b200.sweep import FrequencySweep
    g, t = load_golden("wg_medium")
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("EMB_AUX_HOST", mode)
        sw = FrequencySweep(t, g["er"], g["ur"], golden_bcs(g, t), recycle=0)
        res = sw.run([float(f) for f in g["freqs"]])
        out[mode] = (res.S.copy(), [s["iters"] for s in res.stats], list(sw.aux_dims))
        sw.ctx.close()
    assert out["0"][2] == out["1"][2]
    assert out["0"][1] == out["1"][1]
    assert np.array_equal(out["0"][0], out["1"][0])
