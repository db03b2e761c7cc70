"""CPU: oracle restatement of the Stratton-Chu far field (fem/physics/edm/sc.py) against the unmodified reference's output
(tests/golden/farfield_patch.npz: the demo3 flow on the patch look-alike), and the host-side weighted edge normals."""
from types import SimpleNamespace

import numpy as np

from oracle import ned2_oracle as O
from emerge_b200.farfield import weighted_edge_normals

# the reference sums complex64 values with fastmath; its own round-off is ~1e-6 of the pattern maximum
FF_TOL = 2e-5


def _g():
    import os
    from tests.util import GOLDEN
    return dict(np.load(os.path.join(GOLDEN, "farfield_patch.npz"), allow_pickle=False))


def surface_of(g):
    return SimpleNamespace(areas=g["areas"], normals=g["normals"], tri_to_edge=g["tri_to_edge"], edge_centers=g["edge_centers"],
                           n_tris=g["tri_to_edge"].shape[1])


def test_weighted_normals_match_the_reference_loop():
    g = _g()
    ref = O.weighted_edge_normals(g["areas"], g["normals"], g["tri_to_edge"], g["edge_centers"].shape[1])
    got = weighted_edge_normals(surface_of(g))
    assert got.dtype == np.float32 and np.array_equal(got, ref)


def test_oracle_far_field_matches_reference():
    g = _g()
    wns = O.weighted_edge_normals(g["areas"], g["normals"], g["tri_to_edge"], g["edge_centers"].shape[1])
    E, H = O.stratton_chu_ff(g["Ein"], g["Hin"], g["edge_centers"], wns, g["theta"], g["phi"], float(g["k0"]))
    assert np.abs(E - g["E"]).max() <= FF_TOL * np.abs(g["E"]).max()
    assert np.abs(H - g["H"]).max() <= FF_TOL * np.abs(g["H"]).max()
    assert np.abs(g["E"]).max() > 0
