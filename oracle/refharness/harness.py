"""Drives the UNMODIFIED reference on synthetic meshes.  TEST INFRASTRUCTURE.

Used (a) in the build container to generate the golden fixtures under tests/golden/ (tests/golden/make_golden.py) and
to validate the numpy restatement in oracle/, and (b) on the GPU box by the `-m gpu` drop-in tests, which run the
reference's own frequency_domain() with and without emerge_b200.dropin.install().  The reference package is taken from
oracle/_ref (verbatim copy made by oracle/make_ref.py, git-ignored, travels to the GPU box) or, in the build container
only, from /root/reference.  Recipe: SURVEY.md Appendix C.  Nothing in emerge_b200/ imports this.
"""
from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
_REF_LOCAL = os.path.join(os.path.dirname(HERE), "_ref")
REFERENCE = _REF_LOCAL if os.path.isdir(os.path.join(_REF_LOCAL, "fem")) else "/root/reference"


def setup_paths():
    # the numba cache travels with oracle/_ref so that the GPU box does not have to JIT the whole package again
    cache = os.path.join(_REF_LOCAL, "numba_cache") if REFERENCE == _REF_LOCAL else "/tmp/numba_cache"
    os.environ.setdefault("NUMBA_CACHE_DIR", cache)
    for p in (REFERENCE, os.path.join(HERE, "stubs"), REPO):
        if p not in sys.path:
            sys.path.insert(0, p)
    # stubs must shadow anything else
    sys.path.remove(os.path.join(HERE, "stubs"))
    sys.path.insert(0, os.path.join(HERE, "stubs"))


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE, "fem"))


class _Vol:
    def __init__(self, tag, material):
        self.dimtags = [(3, tag)]
        self.material = material


class FakeMesher:
    """The few attributes Electrodynamics3D/Mesh3D read from fem.mesher.Mesher
    (fem/mesher.py:87-95, fem/mesh3d.py:368-377)."""

    def __init__(self, volumes, boundary_tags):
        self.volumes = volumes
        self.domain_boundary_face_tags = list(boundary_tags)
        self.max_size = None
        self.min_size = None


def build_physics(box, materials=None, pec_extra_tags=()):
    """box: emerge_b200.synthmesh.BoxMesh.  materials: {vol_tag: fem.Material}.
    Returns (fem module, physics, mesh)."""
    setup_paths()
    import gmsh  # the stub
    import fem
    from fem.mesh3d import Mesh3D
    from fem.physics.edm.emfreq3d import Electrodynamics3D
    from fem.solver import ParallelRoutine

    gmsh.set_provider(box)
    if materials is None:
        materials = {1: fem.VACUUM}
    vols = [_Vol(t, m) for t, m in materials.items()]
    boundary = [t for t in sorted(set(box.face_tag.tolist())) if t <= 6] + list(pec_extra_tags)
    mesher = FakeMesher(vols, boundary)
    phys = Electrodynamics3D(mesher)
    mesh = Mesh3D(mesher)
    mesh.update()
    phys.mesh = mesh
    phys._initialize_bcs()                      # PEC on all boundary faces (emfreq3d.py:137-147)
    phys.solveroutine = ParallelRoutine()       # RCM + SuperLU, the reference's own fallback
    return fem, phys, mesh


def rect_waveguide_ports(fem, phys, box, power=1.0):
    """Two RectangularWaveguide ports on z=0 (tag 5) and z=L (tag 6); SURVEY App. A.10 workaround
    (`dims` must be set after construction)."""
    a, b, L = box.dims
    ports = []
    for num, (tag, z, zsign) in enumerate([(5, 0.0, -1.0), (6, L, 1.0)], start=1):
        cs = fem.CoordinateSystem(fem.XAX, fem.YAX, fem.ZAX, origin=__import__("numpy").array([0.0, 0.0, z]))
        port = fem.bc.RectangularWaveguide(fem.FaceSelection([tag]), num, cs=cs, dims=(a, b), power=power)
        port.dims = (a, b)
        ports.append(port)
    phys.assign(*ports)
    return ports
