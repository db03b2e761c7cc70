"""Host-side mirror of the boundary-condition interface the hot path consumes.

The reference's classes live in fem/bc.py; the assembler/sweep only call a handful of members
(`tags`, `port_number`, `active`, `get_gamma(k0)`, `get_Uinc(x,y,k0)`, `get_inv_basis()`, `cs.origin`,
`port_mode_3d_global`, `Zmode`, `modetype`, `v_integration`, `_include_force/_include_stiff`), so reference
objects can be passed to `FrequencySweep` unchanged.  The classes below re-state that interface (same names,
argument meaning, constants) for use where the reference is not importable (GPU box, benchmarks).
They hold triangle ids directly instead of gmsh face tags.
"""
from __future__ import annotations

import numpy as np

C0 = 299792458.0


class CoordSys:
    """Minimal fem.cs.CoordinateSystem: columns of `basis` are the local x,y,z axes (fem/cs.py:240-247)."""

    def __init__(self, xax=(1, 0, 0), yax=(0, 1, 0), zax=(0, 0, 1), origin=(0, 0, 0)):
        ax = [np.asarray(a, dtype=float) / np.linalg.norm(a) for a in (xax, yax, zax)]
        self._basis = np.array(ax).T
        self._basis_inv = np.linalg.pinv(self._basis)
        self.origin = np.asarray(origin, dtype=float)

    def in_local_cs(self, x, y, z):                         # fem/cs.py:370-390
        B = self._basis_inv
        xg, yg, zg = x - self.origin[0], y - self.origin[1], z - self.origin[2]
        return (B[0, 0] * xg + B[0, 1] * yg + B[0, 2] * zg, B[1, 0] * xg + B[1, 1] * yg + B[1, 2] * zg,
                B[2, 0] * xg + B[2, 1] * yg + B[2, 2] * zg)

    def in_global_basis(self, x, y, z):                     # fem/cs.py:392-408
        b = self._basis
        return (b[0, 0] * x + b[0, 1] * y + b[0, 2] * z, b[1, 0] * x + b[1, 1] * y + b[1, 2] * z,
                b[2, 0] * x + b[2, 1] * y + b[2, 2] * z)

    def in_local_basis(self, x, y, z):                      # fem/cs.py:410-430
        B = self._basis_inv
        return (B[0, 0] * x + B[0, 1] * y + B[0, 2] * z, B[1, 0] * x + B[1, 1] * y + B[1, 2] * z,
                B[2, 0] * x + B[2, 1] * y + B[2, 2] * z)


class BoundaryCondition:
    def __init__(self, tri_ids):
        self.tri_ids = np.asarray(tri_ids, dtype=np.int64)


class PEC(BoundaryCondition):
    """fem/bc.py:135-146"""


class RobinBC(BoundaryCondition):
    _include_stiff = False
    _include_mass = False
    _include_force = False
    v_integration = False


class AbsorbingBoundary(RobinBC):
    """First-order ABC, gamma = j k0, no forcing (fem/bc.py:254-301)."""
    _include_stiff = True
    _include_mass = True
    _include_force = False

    def get_gamma(self, k0):
        return 1j * k0

    def get_Uinc(self, x, y, k0):
        return np.zeros((3, len(x)), dtype=np.complex128)


class PortBC(RobinBC):
    Zvac = 376.730313412                                    # fem/bc.py:178
    _include_stiff = True
    _include_force = True
    modetype = "TEM"
    Z0 = None
    power = 1.0

    def __init__(self, tri_ids, port_number, cs: CoordSys, active=False):
        super().__init__(tri_ids)
        self.port_number = port_number
        self.cs = cs
        self.active = active

    def get_inv_basis(self):
        return self.cs._basis_inv

    def get_beta(self, k0):
        return k0

    def get_gamma(self, k0):
        return 1j * self.get_beta(k0)

    def Zmode(self, k0):                                    # fem/bc.py:207-215
        if self.modetype == "TEM":
            return self.Zvac
        if self.modetype == "TE":
            return k0 * 299792458 / self.get_beta(k0) * 4 * np.pi * 1e-7
        return self.get_beta(k0) / (k0 * 299792458 * 8.854187818814 * 1e-12)

    def port_mode_3d_global(self, xg, yg, zg, k0, which="E"):   # fem/bc.py:241-251
        xl, yl, _ = self.cs.in_local_cs(xg, yg, zg)
        Ex, Ey, Ez = self.port_mode_3d(xl, yl, k0)
        return np.array(self.cs.in_global_basis(Ex, Ey, Ez))


class RectangularWaveguide(PortBC):
    """TE_mn rectangular port with analytic mode field (fem/bc.py:496-618).
    Quirk kept for parity: the reference sets `self.type = 'TE'` (fem/bc.py:530) but its `modetype` property is the
    inherited PortBC one and always answers 'TEM' (fem/bc.py:203-205), so S-parameter normalisation uses the TEM
    constants (Zvac, sqrt(eps/mu))."""
    modetype = "TEM"

    def __init__(self, tri_ids, port_number, cs, dims, active=False, power=1.0, mode=(1, 0)):
        super().__init__(tri_ids, port_number, cs, active)
        self.dims = tuple(dims)
        self.power = power
        self.mode = mode

    def get_amplitude(self, k0):
        return np.sqrt(self.power * 4 * 377 / (self.dims[0] * self.dims[1]))     # fem/bc.py:554-555

    def get_beta(self, k0):
        w, h = self.dims
        return np.sqrt(k0 ** 2 - (np.pi * self.mode[0] / w) ** 2 - (np.pi * self.mode[1] / h) ** 2)

    def get_Uinc(self, x, y, k0):
        return -2j * self.get_beta(k0) * self.port_mode_3d(x, y, k0)

    def port_mode_3d(self, x, y, k0, which="E"):
        w, h = self.dims
        E = self.get_amplitude(k0) * np.cos(np.pi * self.mode[0] * x / w) * np.cos(np.pi * self.mode[1] * y / h)
        return np.array([0 * E, E, 0 * E])


class LumpedPort(PortBC):
    """Uniform-field lumped port with sheet impedance (fem/bc.py:620-782)."""
    v_integration = True

    def __init__(self, tri_ids, port_number, cs, width, height, direction, active=False, power=1.0, Z0=50.0):
        super().__init__(tri_ids, port_number, cs, active)
        self.width, self.height = width, height
        self.direction = np.asarray(direction, dtype=float)
        self.power, self.Z0 = power, Z0
        self.vint = None      # (start, end) of the voltage integration line, set by the sweep

    @property
    def surfZ(self):
        return self.Z0 * self.width / self.height

    @property
    def voltage(self):
        return np.sqrt(2 * self.power * self.Z0)

    def get_gamma(self, k0):
        return 1j * k0 * 376.7303 / self.surfZ

    def get_Uinc(self, x, y, k0):
        Emag = -1j * 2 * k0 * self.voltage / self.height * (376.7303 / self.surfZ)
        return Emag * self.port_mode_3d(x, y, k0)

    def port_mode_3d(self, x, y, k0, which="E"):
        px, py, pz = self.cs.in_local_basis(*self.direction)
        o = np.ones_like(x)
        return np.array([px * o, py * o, pz * o])


class SampledField:
    """Vector field known at a fixed set of points (3, n): the port-mode field of a boundary-mode analysis sampled at the
    Dunavant points of the port triangles - the only points the hot path ever evaluates it at (assembler.py:63-98,
    mth/sparam.py:72-139).  Lookup is by nearest stored point, so the local->global round trip of ModalPort.port_mode_3d
    (fem/bc.py:469-480) lands on the same sample."""

    def __init__(self, pts, values):
        from scipy.spatial import cKDTree
        self.pts = np.ascontiguousarray(np.asarray(pts, dtype=float).T)
        self.values = np.asarray(values, dtype=np.complex128)
        self.tree = cKDTree(self.pts)
        span = np.ptp(self.pts, axis=0).max()
        self.tol = 1e-6 * (span if span > 0 else 1.0)

    def __call__(self, x, y, z):
        q = np.stack([np.ravel(x), np.ravel(y), np.ravel(z)], axis=1)
        d, i = self.tree.query(q)
        if d.size and d.max() > self.tol:
            raise ValueError("SampledField: evaluation point is not one of the stored sample points")
        return self.values[:, i]


class ModalPort(PortBC):
    """Port whose mode field comes from a boundary-mode analysis done upstream (fem/bc.py:329-494).
    `E_function(xg,yg,zg) -> (3,n)` is the mode field in global coordinates (already normalised)."""

    def __init__(self, tri_ids, port_number, cs, E_function, beta, k0_mode, TEM=True, freq_mode=None,
                 active=False, power=1.0, modetype="TEM"):
        super().__init__(tri_ids, port_number, cs, active)
        self.E_function, self.beta0, self.k0_mode = E_function, beta, k0_mode
        self.TEM, self.freq_mode, self.power = TEM, freq_mode, power
        self.modetype = modetype

    def get_beta(self, k0):                                 # fem/bc.py:434-446
        if self.TEM:
            return self.beta0 / self.k0_mode * k0
        freq = k0 * 299792458 / (2 * np.pi)
        return np.sqrt(self.beta0 ** 2 + k0 ** 2 * (1 - ((self.freq_mode / freq) ** 2)))

    def get_Uinc(self, x, y, k0):
        return -2j * self.get_beta(k0) * self.port_mode_3d(x, y, k0)

    def port_mode_3d(self, x, y, k0, which="E"):            # fem/bc.py:469-480
        b = self.cs._basis
        xg = b[0, 0] * x + b[0, 1] * y + self.cs.origin[0]
        yg = b[1, 0] * x + b[1, 1] * y + self.cs.origin[1]
        zg = b[2, 0] * x + b[2, 1] * y + self.cs.origin[2]
        Eg = self.port_mode_3d_global(xg, yg, zg, k0)
        return np.array(self.cs.in_local_basis(Eg[0], Eg[1], Eg[2]))

    def port_mode_3d_global(self, xg, yg, zg, k0, which="E"):
        return np.sqrt(self.power) * np.asarray(self.E_function(xg, yg, zg))
