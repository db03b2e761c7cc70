"""CPU: oracle restatement of the field post-processing (point location, E and curl interpolation) against values
produced by the unmodified reference's EMDataSet.interpolate (tests/golden/interp_wg_tiny.npz)."""
import numpy as np

from oracle import ned2_oracle as O
from tests.util import load_golden


def test_locate_and_interpolate_match_reference():
    g, t = load_golden("interp_wg_tiny")
    pts = g["pts"]
    tet_of = O.locate_points(t.nodes, t.tets, pts)
    outside = np.abs(g["E"]).sum(axis=0) == 0
    assert np.array_equal(tet_of < 0, outside)                      # same points fall outside the mesh
    const = 1.0 / (-1j * 2 * np.pi * float(g["freq"]) * (g["ur00"] * 4 * np.pi * 1e-7))      # emdata.py:193
    E, H = O.interp_fields(t.nodes, t.tets, t.edges, t.tris, t.tet_to_field, g["x"], pts, tet_of, const)
    assert np.abs(E - g["E"]).max() <= 1e-10 * np.abs(g["E"]).max()
    assert np.abs(H - g["H"]).max() <= 1e-10 * np.abs(g["H"]).max()
