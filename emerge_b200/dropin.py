"""Drop-in replacements for the two seams the reference exposes on its frequency-domain path:

  * `physics.assembler`     -> GpuAssembler      : Assembler.assemble_freq_matrix (fem/physics/edm/assembler.py:312-388)
  * `physics.solveroutine`  -> solve() patched   : SolveRoutine.solve             (fem/solver.py:405-469)

Same method names, argument meaning, return shapes and error behaviour; the work behind them runs through the C ABI
(include/emerge_b200.h) on the GPU.  `install(physics)` wires both into a reference `Electrodynamics3D`, after which
the reference's own `frequency_domain()` (emfreq3d.py:607-732) runs unchanged on top of the CUDA path.

What differs, deliberately:
  * K/M, the surface matrices, the eliminated pattern and the auxiliary spaces are built once (first call); per frequency
    only gamma_p(f) and the incident-field samples cross the boundary (the reference re-assembles every surface term and
    rebuilds solve_ids in Python at every frequency, assembler.py:333-385).
  * The returned K is a `DeviceCSR`: a scipy csr_matrix subclass that holds a handle to A(f) on the device and only
    downloads values when a caller actually reads them (`.materialize()`; 5.3 GB at 1M tets).  Its rows/columns of
    eliminated (PEC) dofs are empty - the reference never reads them after `A[np.ix_(solve_ids, solve_ids)]`
    (solver.py:434).
  * port_vectors are `DeviceRHS` ndarrays carrying the id of the device-resident right-hand side, so
    `b + port_vectors[p]` (emfreq3d.py:691) followed by `solveroutine.solve` never uploads a length-N vector.
  * Non-convergence raises `emerge_b200.lib.NotConverged` instead of being logged and ignored (solver.py:467-468).
There is no CPU fallback: without the CUDA library or a GPU, constructing GpuAssembler raises.
"""
from __future__ import annotations

import types

import numpy as np
import scipy.sparse as sp

from .lib import Context, EmergeB200Error
from .sweep import FrequencySweep


class _RefTables:
    """The arrays FrequencySweep reads, taken verbatim from the reference's Nedelec2 field and its Mesh3D
    (fem/elements/nedelec2.py:46-62, fem/mesh3d.py:227-352).  Never renumbered."""

    def __init__(self, field):
        mesh = field.mesh
        for name in ("nodes", "tets", "edges", "tris", "tet_to_edge", "tet_to_tri", "tri_to_edge", "tri_to_tet",
                     "edge_lengths"):
            setattr(self, name, np.asarray(getattr(mesh, name)))
        for name in ("tet_to_field", "tri_to_field", "edge_to_field"):
            setattr(self, name, np.asarray(getattr(field, name)))

    @property
    def n_field(self):
        return 2 * self.edges.shape[1] + 2 * self.tris.shape[1]


class DeviceRHS(np.ndarray):
    """Length-N right-hand side whose values also live on the device as the forcing of surface `sid`.
    Adding the all-zero `b` of assemble_freq_matrix keeps the tag; any other arithmetic drops it (the solve then uploads
    the host values)."""

    def __new__(cls, values, sid, owner):
        obj = np.asarray(values, dtype=np.complex128).view(cls)
        obj._emb_sid, obj._emb_owner = sid, owner
        return obj

    def __array_finalize__(self, obj):
        self._emb_sid = getattr(obj, "_emb_sid", None)
        self._emb_owner = getattr(obj, "_emb_owner", None)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        raw = [np.asarray(x) if isinstance(x, np.ndarray) else x for x in inputs]
        out = getattr(ufunc, method)(*raw, **kwargs)
        tagged = [x for x in inputs if isinstance(x, DeviceRHS) and x._emb_sid is not None]
        if (ufunc is np.add and method == "__call__" and len(tagged) == 1 and isinstance(out, np.ndarray)
                and all((x is tagged[0]) or (isinstance(x, np.ndarray) and getattr(x, "_emb_zero", False) and not x.any())
                        for x in inputs)):
            return DeviceRHS(out, tagged[0]._emb_sid, tagged[0]._emb_owner)
        return out


class _ZeroRHS(np.ndarray):
    """the all-zero `b` of assemble_freq_matrix (assembler.py:343); tagged so that b + port_vector stays on the device"""

    def __new__(cls, n):
        obj = np.zeros(n, dtype=np.complex128).view(cls)
        obj._emb_zero = True
        return obj

    def __array_finalize__(self, obj):
        self._emb_zero = False        # results of arithmetic are ordinary arrays; only __new__ sets the tag


class DeviceCSR(sp.csr_matrix):
    """scipy csr_matrix (N x N, complex128) standing for A(f) on the device.  Holds no values until materialize()."""

    @classmethod
    def lazy(cls, n, owner, freq):
        m = cls((n, n), dtype=np.complex128)
        m._emb_owner, m._emb_freq, m._emb_loaded = owner, float(freq), False
        return m

    def materialize(self):
        """downloads A(f) (solve-space values expanded to N x N) into this matrix and returns it"""
        own = self._emb_owner
        if own is None or self._emb_loaded:
            return self
        if own._current_freq != self._emb_freq:
            raise EmergeB200Error("DeviceCSR: the device now holds A(f) of another frequency; materialize() it before the "
                                  "next assemble_freq_matrix call")
        ip, ix, data = own.sweep.ctx.get_csr(2)
        sid = own.solve_ids
        ns = len(sid)
        full = sp.csr_matrix((data, ix, ip), shape=(ns, ns)).tocoo()
        out = sp.csr_matrix((full.data, (sid[full.row], sid[full.col])), shape=self.shape)
        self.data, self.indices, self.indptr = out.data, out.indices, out.indptr
        self._emb_loaded = True
        return self


class GpuAssembler:
    """Replacement for fem.physics.edm.assembler.Assembler on the frequency-domain path."""

    def __init__(self, device: int = 0, recycle: int = 40, multilevel: bool = True, rtol: float = 1e-8,
                 solver_opts: dict | None = None, f_ref: float | None = None, ctx=None):
        self.cached_matrices = None           # Assembler.cached_matrices (assembler.py:208): True-ish once K/M are on the device
        self.device, self.recycle, self.multilevel = device, recycle, multilevel
        self.solver_opts = dict(rtol=rtol, **(solver_opts or {}))
        self.f_ref = f_ref
        self.sweep: FrequencySweep | None = None
        self.solve_ids = None
        self._current_freq = None
        self._key = None
        self.ctx = ctx if ctx is not None else Context(device)   # fails loudly here without the CUDA library / a GPU

    def _bind(self, field, er, ur, bcs, frequency):
        key = (id(field), id(er), id(ur), tuple(id(b) for b in bcs))
        if self.sweep is not None and key == self._key:
            return
        mesh = field.mesh
        self.sweep = FrequencySweep(_RefTables(field), er, ur, bcs, device=self.device,
                                    get_triangles=getattr(mesh, "get_triangles", None), ctx=self.ctx,
                                    recycle=self.recycle, multilevel=self.multilevel,
                                    f_ref=self.f_ref if self.f_ref is not None else float(frequency))
        self.sweep.solver_opts.update(self.solver_opts)
        self.sweep.setup()
        self.solve_ids = self.ctx.solve_ids()
        self._key = key
        self.cached_matrices = ("device", "device")

    def assemble_freq_matrix(self, field, er, ur, bcs, frequency, cache_matrices=False):
        """-> (K, b, solve_ids, port_vectors) as assembler.py:312-388.  cache_matrices is accepted for signature
        compatibility; K/M always stay cached on the device for the lifetime of (field, er, ur, bcs)."""
        self._bind(field, er, ur, bcs, frequency)
        sw = self.sweep
        n = sw.ctx.n_field
        k0 = 2 * np.pi * frequency / 299792458
        sids, gammas = [], []
        port_vectors = {}
        for bc in sw.robin:
            sid = sw.sid[id(bc)]
            if bc._include_force:
                xy = sw.points[id(bc)]
                U = np.asarray(bc.get_Uinc(xy[0].ravel(), xy[1].ravel(), k0), dtype=np.complex128)
                full = sw.ctx.surface_set_U(sid, U.reshape(3, 6, sw.ntri[id(bc)]), want_full=True)
                port_vectors[bc.port_number] = DeviceRHS(full, sid, self)
            if bc._include_stiff:
                sids.append(sid)
                gammas.append(complex(bc.get_gamma(k0)))
        for bc in sw.ports:
            port_vectors.setdefault(bc.port_number, np.zeros(n, dtype=np.complex128))
        sw.ctx.form_A(k0, sids, gammas)
        self._current_freq = float(frequency)
        return DeviceCSR.lazy(n, self, frequency), _ZeroRHS(n), self.solve_ids.copy(), port_vectors

    # ---- solver seam ---------------------------------------------------------------------------------------------
    def solve(self, A, b, solve_ids, reuse=False):
        """SolveRoutine.solve(A, b, solve_ids, reuse) -> x (length N, zeros at eliminated dofs), solver.py:405-469.
        A must be the DeviceCSR of the last assemble_freq_matrix call (the device holds exactly that A(f))."""
        if getattr(A, "_emb_owner", None) is not self or A._emb_freq != self._current_freq:
            raise EmergeB200Error("GpuAssembler.solve: A is not the matrix of the last assemble_freq_matrix call; "
                                  "this path has no CPU fallback for foreign matrices")
        if len(solve_ids) != len(self.solve_ids):
            raise EmergeB200Error("solve_ids differ from the eliminated pattern on the device")
        sid = getattr(b, "_emb_sid", None)
        if sid is not None and getattr(b, "_emb_owner", None) is self:
            x, info = self.ctx.solve(sid, **self.sweep.solver_opts)
        else:
            x, info = self.ctx.solve_rhs(np.asarray(b), **self.sweep.solver_opts)
        self.last_info = info
        return x


def install(physics, **kw) -> GpuAssembler:
    """Wires the CUDA path into a reference Electrodynamics3D: replaces `physics.assembler` (emfreq3d.py:94) and the
    `solve` method of `physics.solveroutine` (emfreq3d.py:98); `solveroutine.eig` (modal analysis) is left untouched."""
    asm = GpuAssembler(**kw)
    physics.assembler = asm
    routine = physics.solveroutine

    def solve(self, A, b, solve_ids, reuse=False):
        return asm.solve(A, b, solve_ids, reuse)
    routine.solve = types.MethodType(solve, routine)
    return asm
