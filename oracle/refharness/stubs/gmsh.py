"""Fake `gmsh` module serving a synthetic mesh to the unmodified reference (TEST INFRASTRUCTURE).

Only the calls the reference's hot path makes are provided (fem/mesh3d.py:162,185,198,217,225,
236,315,372; fem/selection.py:197-200; fem/simmodel.py:193-199).  A provider object holding
numpy arrays is installed with `set_provider`.
"""
import numpy as np

_P = None


def set_provider(p):
    global _P
    _P = p


def initialize(*a, **k):
    pass


def finalize(*a, **k):
    pass


def is_initialized():
    return True


class _Mesh:
    @staticmethod
    def get_nodes(*a, **k):
        n = _P.nodes_xyz.shape[0]
        return np.arange(1, n + 1), _P.nodes_xyz.reshape(-1).copy(), None

    @staticmethod
    def get_elements(dim=-1, tag=-1):
        if dim == 3:
            if tag is None or tag < 0:
                ids = np.arange(_P.tets.shape[0])
            else:
                ids = np.nonzero(_P.tet_vol == tag)[0]
            return [4], [ids + 1], [(_P.tets[ids] + 1).reshape(-1)]
        if dim == 2:
            if tag is None or tag < 0:
                ids = np.arange(_P.face_tris.shape[0])
            else:
                ids = np.nonzero(_P.face_tag == tag)[0]
            return [2], [ids + 1], [(_P.face_tris[ids] + 1).reshape(-1)]
        raise NotImplementedError(dim)

    getNodes = get_nodes
    getElements = get_elements


class _Model:
    mesh = _Mesh()

    @staticmethod
    def getNormal(tag, uv):
        return np.asarray(_P.face_normal[tag], dtype=float)

    get_normal = getNormal

    @staticmethod
    def get_boundary(dimtags, *a, **k):
        return [(2, t) for t in sorted(set(_P.face_tag.tolist()))]


model = _Model()


class option:
    @staticmethod
    def setNumber(*a, **k):
        pass

    set_number = setNumber
