"""CPU: Touchstone writer / reader round trip (2-port column order, n-port line wrapping, all three data formats)."""
import numpy as np
import pytest

from emerge_b200.sweep import SweepResult
from emerge_b200.touchstone import export_touchstone, read_touchstone, write_touchstone


@pytest.mark.parametrize("n", [1, 2, 3, 5])
@pytest.mark.parametrize("fmt", ["RI", "MA", "DB"])
def test_round_trip(tmp_path, n, fmt):
    rng = np.random.default_rng(n)
    f = np.linspace(8e9, 12e9, 7)
    S = rng.standard_normal((7, n, n)) + 1j * rng.standard_normal((7, n, n))
    path = write_touchstone(str(tmp_path / "x"), f, S, fmt, comments=["emerge_b200"])
    assert path.endswith(f".s{n}p")
    f2, S2, z0 = read_touchstone(path)
    assert z0 == 50.0 and np.allclose(f2, f, rtol=1e-12)
    assert np.allclose(S2, S, rtol=1e-10, atol=1e-12)


def test_two_port_column_order_and_result_objects(tmp_path):
    S = np.zeros((1, 2, 2), complex)
    S[0, 1, 0] = 0.5          # S21
    path = export_touchstone(SweepResult(np.array([1e9]), [1, 2], S), str(tmp_path / "y.s2p"))
    row = [float(v) for v in open(path).read().splitlines()[-1].split()]
    assert row[3] == 0.5 and row[5] == 0.0          # f, S11 (re, im), S21 (re, im), S12, S22

    class _Sp:
        map = {1: 0, 2: 1}

        def __call__(self, i, j):
            return S[0, i - 1, j - 1]

    class _Set:
        freq, Sp = 1e9, _Sp()

    class _Data:
        datasets = [_Set()]
    f2, S2, _ = read_touchstone(export_touchstone(_Data(), str(tmp_path / "z"), "MA"))
    assert np.allclose(S2, S)


def test_bad_shapes_raise(tmp_path):
    with pytest.raises(ValueError):
        write_touchstone(str(tmp_path / "a"), np.ones(3), np.ones((3, 2, 3)))
    with pytest.raises(ValueError):
        write_touchstone(str(tmp_path / "a"), np.ones(2), np.ones((3, 2, 2)))
