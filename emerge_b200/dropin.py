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
from .sweep import FrequencySweep, _is


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

    def __new__(cls, values, sid, owner, freq=None):
        obj = np.asarray(values, dtype=np.complex128).view(cls)
        obj._emb_sid, obj._emb_owner, obj._emb_freq = sid, owner, freq
        return obj

    def __array_finalize__(self, obj):
        self._emb_sid = getattr(obj, "_emb_sid", None)
        self._emb_owner = getattr(obj, "_emb_owner", None)
        self._emb_freq = getattr(obj, "_emb_freq", None)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        raw = [np.asarray(x) if isinstance(x, np.ndarray) else x for x in inputs]
        outs = kwargs.get("out")
        if outs is not None:
            # in-place arithmetic (pv += ...): compute on the base arrays, the result no longer matches the device copy
            kwargs["out"] = tuple(np.asarray(o) if isinstance(o, np.ndarray) else o for o in outs)
            getattr(ufunc, method)(*raw, **kwargs)
            for o in outs:
                if isinstance(o, DeviceRHS):
                    o._emb_sid = None
            return outs[0] if len(outs) == 1 else outs
        out = getattr(ufunc, method)(*raw, **kwargs)
        tagged = [x for x in inputs if isinstance(x, DeviceRHS) and x._emb_sid is not None]
        if (ufunc is np.add and method == "__call__" and len(tagged) == 1 and isinstance(out, np.ndarray)
                and all((x is tagged[0]) or (isinstance(x, np.ndarray) and getattr(x, "_emb_zero", False) and not x.any())
                        for x in inputs)):
            return DeviceRHS(out, tagged[0]._emb_sid, tagged[0]._emb_owner, tagged[0]._emb_freq)
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
        self._held = None
        self.ctx = ctx if ctx is not None else Context(device)   # fails loudly here without the CUDA library / a GPU

    def invalidate(self):
        """forget the cached device state (mesh, K/M, surfaces, recycled basis); the next call re-binds"""
        self.sweep, self._held, self.cached_matrices = None, None, None
        self._current_freq = None

    def _same_problem(self, field, er, ur, bcs):
        """True if (field, er, ur, bcs) describe the problem the device state was built for.  Same objects as last time
        (held by strong references, so CPython cannot have recycled their ids): no checks - in-place edits need
        invalidate().  Different objects - the reference rebuilds er / ur on every frequency_domain() call
        (emfreq3d.py:612-616) and a re-mesh or a material sweep produces new ones - are compared by CONTENT, once."""
        if self.sweep is None or self._held is None:
            return False
        hf, he, hu, hb = self._held
        if len(hb) != len(bcs) or any(a is not b for a, b in zip(hb, bcs)):
            return False
        if hf is field and he is er and hu is ur:
            return True
        if hf is not field:
            fm, hm = field.mesh, hf.mesh
            for name in ("tets", "nodes", "edges", "tris"):
                if not np.array_equal(np.asarray(getattr(fm, name)), np.asarray(getattr(hm, name))):
                    return False
            if not np.array_equal(np.asarray(field.tet_to_field), np.asarray(hf.tet_to_field)):
                return False
        return np.array_equal(er, he) and np.array_equal(ur, hu)

    def _bind(self, field, er, ur, bcs, frequency):
        if self._same_problem(field, er, ur, bcs):
            self._held = (field, er, ur, list(bcs))
            return
        self._held = (field, er, ur, list(bcs))     # strong references: CPython may not reuse these ids while cached
        mesh = field.mesh
        self.sweep = FrequencySweep(_RefTables(field), er, ur, bcs, device=self.device,
                                    get_triangles=getattr(mesh, "get_triangles", None), ctx=self.ctx,
                                    recycle=self.recycle, multilevel=self.multilevel,
                                    f_ref=self.f_ref if self.f_ref is not None else float(frequency))
        self.sweep.solver_opts.update(self.solver_opts)
        self.sweep.setup()
        self.solve_ids = self.ctx.solve_ids()
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
                port_vectors[bc.port_number] = DeviceRHS(full, sid, self, float(frequency))
            if bc._include_stiff:
                sids.append(sid)
                gammas.append(complex(bc.get_gamma(k0)))
        for bc in sw.ports:
            port_vectors.setdefault(bc.port_number, np.zeros(n, dtype=np.complex128))
        sw.ctx.form_A(k0, sids, gammas)
        self._current_freq = float(frequency)
        return DeviceCSR.lazy(n, self, frequency), _ZeroRHS(n), self.solve_ids.copy(), port_vectors

    # ---- solver seam ---------------------------------------------------------------------------------------------
    def assemble_bma_matrices(self, field, er, ur, k0, port, bcs):
        """Assembler.assemble_bma_matrices (assembler.py:246-308): matrices of the port's boundary-mode eigenproblem.
        Same arguments and return value (E, B, solve_ids, NedelecLegrange2 field).  The element loop - the numba prange
        kernel generalized_matrix_GQ, nedeleclegrange2.py:223-417 - runs on the device (emb_bma_element_matrices); the
        port's SurfaceMesh / NedelecLegrange2 objects and the PEC bookkeeping are the reference's own, because
        modal_analysis keeps using them afterwards (emfreq3d.py:330-362)."""
        import importlib
        pkg = type(field).__module__.rsplit(".", 1)[0]                       # "...fem.elements"
        NedelecLegrange2 = importlib.import_module(pkg + ".nedleg2").NedelecLegrange2
        mesh = field.mesh
        tri_ids = mesh.get_triangles(port.tags)
        origin = tuple([c - n for c, n in zip(port.cs.origin, port.cs.gzhat)])
        surf = mesh.boundary_surface(port.tags, origin)
        nlf = NedelecLegrange2(surf, port.cs)
        xy = (np.linalg.pinv(port.cs._basis) @ surf.nodes)[:2]               # nedeleclegrange2.py:43
        eA, eB = self.ctx.bma_element_matrices(xy, surf.tris, surf.edges, nlf.tri_to_field[:3, :], er[:, :, tri_ids],
                                               ur[:, :, tri_ids], k0)
        ttf = np.asarray(nlf.tri_to_field, dtype=np.int64)
        rows = np.repeat(ttf.T[:, :, None], 14, axis=2).ravel()
        cols = np.repeat(ttf.T[:, None, :], 14, axis=1).ravel()
        n = nlf.n_field
        E = sp.coo_matrix((eA.ravel(), (rows, cols)), shape=(n, n)).tocsr()
        B = sp.coo_matrix((eB.ravel(), (rows, cols)), shape=(n, n)).tocsr()
        pec_ids, pec_edges, pec_vertices = [], [], []
        for bc in bcs:                                                       # assembler.py:268-296
            if not _is(bc, "PEC"):
                continue
            tids = mesh.get_triangles(bc.tags)
            for ii in list(mesh.tri_to_edge[:, tids].flatten()):
                i2 = surf.from_source_edge(ii)
                if i2 is None:
                    continue
                eids = nlf.edge_to_field[:, i2]
                pec_ids.extend(list(eids))
                pec_edges.append(eids[0])
                pec_vertices.append(eids[3] - nlf.n_xy)
                pec_vertices.append(eids[4] - nlf.n_xy)
            for ii in tids:
                i2 = surf.from_source_tri(ii)
                if i2 is None:
                    continue
                pec_ids.extend(list(nlf.tri_to_field[:, i2]))
        port._field, port._pece, port._pecv = nlf, pec_edges, pec_vertices
        gone = np.zeros(n, dtype=bool)
        if pec_ids:
            gone[np.asarray(pec_ids, dtype=np.int64)] = True
        return E, B, np.nonzero(~gone)[0], nlf

    def solve(self, A, b, solve_ids, reuse=False):
        """SolveRoutine.solve(A, b, solve_ids, reuse) -> x (length N, zeros at eliminated dofs), solver.py:405-469.
        A must be the DeviceCSR of the last assemble_freq_matrix call (the device holds exactly that A(f))."""
        if getattr(A, "_emb_owner", None) is not self or A._emb_freq != self._current_freq:
            raise EmergeB200Error("GpuAssembler.solve: A is not the matrix of the last assemble_freq_matrix call; "
                                  "this path has no CPU fallback for foreign matrices")
        if len(solve_ids) != len(self.solve_ids):
            raise EmergeB200Error("solve_ids differ from the eliminated pattern on the device")
        sid = getattr(b, "_emb_sid", None)
        # the device-resident right-hand side is only valid for the frequency it was computed at
        if sid is not None and getattr(b, "_emb_owner", None) is self and getattr(b, "_emb_freq", None) == self._current_freq:
            x, info = self.ctx.solve(sid, **self.sweep.solver_opts)
        else:
            x, info = self.ctx.solve_rhs(np.asarray(b), **self.sweep.solver_opts)
        self.last_info = info
        return x


def _sim_error(physics):
    import importlib
    return getattr(importlib.import_module(type(physics).__module__), "SimulationError", RuntimeError)


def install_postproc(physics, asm: GpuAssembler):
    """Routes EMDataSet.interpolate (emdata.py:181-199) through the device: replaces `basis.interpolate` and
    `basis.interpolate_curl` (fem/elements/nedelec2.py:72-86) of the physics' basis by versions that locate the points and
    evaluate E / curl E with emb_interp_fields.  Calls that restrict the candidate tets (`tetids=...`, the S-parameter
    path of the reference's own sweep loop) keep the reference implementation."""
    basis = physics.basis
    if basis is None or getattr(basis, "_emb_postproc", False):
        return
    ref_interp, ref_curl = basis.interpolate, basis.interpolate_curl

    def interpolate(field, xs, ys, zs, tetids=None):
        if tetids is not None or asm.sweep is None:
            return ref_interp(field, xs, ys, zs, tetids) if tetids is not None else ref_interp(field, xs, ys, zs)
        E, _ = asm.ctx.interp_fields(field, np.array([xs, ys, zs], dtype=float))
        return E[0], E[1], E[2]

    def interpolate_curl(field, xs, ys, zs, c, tetids=None):
        if tetids is not None or asm.sweep is None:
            return ref_curl(field, xs, ys, zs, c, tetids) if tetids is not None else ref_curl(field, xs, ys, zs, c)
        _, H = asm.ctx.interp_fields(field, np.array([xs, ys, zs], dtype=float), curl_const=c)
        return H[0], H[1], H[2]
    basis.interpolate, basis.interpolate_curl = interpolate, interpolate_curl
    basis._emb_postproc = True


def _gpu_frequency_domain(physics, asm: GpuAssembler, dist=None, keep_fields: bool = True):
    """Fast driver behind physics.frequency_domain() / frequency_domain_par(): same preamble, same EMSimData filling as
    emfreq3d.py:607-732 (and :545-601 for the parallel variant), but the per-frequency loop is FrequencySweep's - all ports
    of a point solved in lockstep, recycled reduced basis across points, S-parameters from device-side field samples - so
    no N x N matrix and no length-N vector crosses the host boundary except the solved fields the result object keeps.
    dist: an initialised torch.distributed module -> the points are sharded over the ranks (ShardedSweep); every rank
    returns the complete S-parameters, fields only of its own points."""
    import importlib
    mod = importlib.import_module(type(physics).__module__)
    if physics._bc_initialized is False:
        raise mod.SimulationError("Cannot run a modal analysis because no boundary conditions have been assigned.")
    physics._initialize_field()
    physics._initialize_bc_data()
    mesh = physics.mesh
    er = mesh.retreive(lambda mat, x, y, z: mat.fer3d_mat(x, y, z), physics.mesher.volumes)
    ur = mesh.retreive(lambda mat, x, y, z: mat.fur3d_mat(x, y, z), physics.mesher.volumes)
    physics.data = mod.EMSimData(physics.basis)
    bcs = physics.boundary_conditions
    freqs = list(physics.frequencies)
    asm._bind(physics.basis, er, ur, bcs, float(np.median(freqs)))
    install_postproc(physics, asm)
    sw = asm.sweep
    all_ports = [bc for bc in bcs if isinstance(bc, mod.PortBC)]
    port_numbers = [p.port_number for p in all_ports]
    assert [p.port_number for p in sw.ports] == port_numbers
    for p in all_ports:
        p.active = False
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    if world > 1:
        from .distributed import ShardedSweep
        sh = ShardedSweep(sw, freqs, rank, world, dist=dist, device=asm.device)
        fields = {}
        res = sh.run(keep_fields=keep_fields, fields_out=fields)
        S = sh.gather_S(res)
    else:
        res = sw.run(freqs, keep_fields=keep_fields)
        S, fields = res.S, res.fields
    asm.last_stats = res.stats
    er00, ur00 = np.squeeze(er[0, 0, :]), np.squeeze(ur[0, 0, :])
    for i, freq in enumerate(freqs):
        k0 = 2 * np.pi * freq / 299792458
        data = physics.data.new(freq=freq, k0=k0)
        data.init_sp(port_numbers)
        data.er, data.ur = er00, ur00
        for port in all_ports:
            data.add_port_properties(port.port_number, mode_number=port.mode_number, k0=k0, beta=port.get_beta(k0),
                                     Z0=port.Z0, Pout=port.power)
        for ja, pa in enumerate(all_ports):
            if (i, pa.port_number) in fields:
                data._fields[pa.port_number] = fields[(i, pa.port_number)]
            for ib, pb in enumerate(all_ports):
                data.write_S(pb.port_number, pa.port_number, S[i, ib, ja])
        if data._fields:
            data.set_field_vector()
        else:       # set_field_vector without fields (emdata.py:134-136): excitation of the first port
            data.excitation = {n: 0.0 for n in port_numbers}
            data.excitation[port_numbers[0]] = 1.0 + 0j
    return physics.data


def install(physics, fast: bool = False, keep_fields: bool = True, **kw) -> GpuAssembler:
    """Wires the CUDA path into a reference Electrodynamics3D.

    Always: replaces `physics.assembler` (emfreq3d.py:94) and the `solve` method of `physics.solveroutine`
    (emfreq3d.py:98), so the reference's own frequency_domain() loop runs unchanged on top of the CUDA path; and
    `solveroutine.eig` (solver.py:471-505), so that modal_analysis (emfreq3d.py:201-364) runs its element loop
    (GpuAssembler.assemble_bma_matrices) and the spectral transformation of its eigen-solve (modal.gpu_eig) on the device.

    fast=True additionally replaces the two sweep drivers themselves:
      physics.frequency_domain()          -> EMSimData   (emfreq3d.py:607-732)
      physics.frequency_domain_par(njobs) -> EMSimData   (emfreq3d.py:469-605)
    by FrequencySweep-backed versions that fill the same EMSimData (Sp, _fields[port], er/ur[0,0,:], port properties);
    the result object's .interpolate(xs, ys, zs) (emdata.py:181-199) then evaluates E and H on the device (install_postproc)
    and `fem.physics.edm.stratton_chu` (sc.py:144-180) runs on the device (farfield.install_farfield).
    frequency_domain_par shards the frequency points over the ranks of the initialised torch.distributed process group
    (one process per GPU, launched with torchrun - the counterpart of the reference's multiprocessing.Pool(njobs)); njobs
    is accepted for signature compatibility, the parallel width is the world size.  Without a process group it runs on
    this process's GPU."""
    asm = GpuAssembler(**kw)
    physics.assembler = asm
    routine = physics.solveroutine

    def solve(self, A, b, solve_ids, reuse=False):
        return asm.solve(A, b, solve_ids, reuse)
    routine.solve = types.MethodType(solve, routine)
    if hasattr(routine, "eig"):          # modal_analysis: element loop via asm.assemble_bma_matrices, eigen-solve operator on the device
        from .modal import install_modal
        install_modal(physics, asm)
    if fast:
        def frequency_domain(self):
            return _gpu_frequency_domain(self, asm, None, keep_fields)

        def frequency_domain_par(self, njobs: int = 2):
            dist = None
            try:
                import torch.distributed as td
                if td.is_available() and td.is_initialized():
                    dist = td
            except ImportError:
                pass
            return _gpu_frequency_domain(self, asm, dist, keep_fields)
        physics.frequency_domain = types.MethodType(frequency_domain, physics)
        physics.frequency_domain_par = types.MethodType(frequency_domain_par, physics)
        # far field: fem.physics.edm.stratton_chu (sc.py:144-180) -> emb_stratton_chu on the assembler's context
        import importlib
        pkg = type(physics).__module__.rsplit(".", 1)[0]            # "...physics.edm"
        try:
            from .farfield import install_farfield
            install_farfield(importlib.import_module(pkg), lambda: asm.ctx)
        except ImportError:
            pass
    return asm
