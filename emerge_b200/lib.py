"""ctypes binding of the C ABI in include/emerge_b200.h (libemerge_b200.so, CUDA sm_100a).

There is no CPU fallback: if the shared library is missing or no GPU is visible, constructing a
`Context` raises.  Loading the library itself (symbol checks) works without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libemerge_b200.so")

c128_p = C.c_void_p


class SolveOpts(C.Structure):
    _fields_ = [("method", C.c_int), ("precond", C.c_int), ("restart", C.c_int), ("maxit", C.c_int),
                ("rtol", C.c_double), ("use_x0", C.c_int)]


class SolveInfo(C.Structure):
    _fields_ = [("iters", C.c_int), ("relres", C.c_double), ("ms", C.c_double), ("spmvs", C.c_int)]


METHODS = {"gmres": 0, "bicgstab": 1, "cocr": 2}
PRECONDS = {"none": 0, "jacobi": 1, "block": 2, "multilevel": 3}

_SIGS = {
    "emb_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "emb_destroy": (None, [C.c_void_p]),
    "emb_last_error": (C.c_char_p, [C.c_void_p]),
    "emb_version": (C.c_char_p, []),
    "emb_launch_count": (C.c_int64, [C.c_void_p]),
    "emb_last_ms": (C.c_double, [C.c_void_p, C.c_char_p]),
    "emb_timer_start": (C.c_int, [C.c_void_p]),
    "emb_timer_stop": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "emb_profiler": (C.c_int, [C.c_void_p, C.c_int]),
    "emb_upload_mesh": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64] + [C.c_void_p] * 5),
    "emb_upload_materials": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "emb_locate_points": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "emb_interp_fields": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "emb_stratton_chu": (C.c_int, [C.c_void_p, C.c_int64] + [C.c_void_p] * 4 + [C.c_int64, C.c_void_p, C.c_void_p, C.c_double,
                                   C.c_void_p, C.c_void_p]),
    "emb_bma_element_matrices": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64] + [C.c_void_p] * 6 + [C.c_double, C.c_void_p,
                                           C.c_void_p]),
    "emb_shift_invert_setup": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_double, C.c_double]),
    "emb_shift_invert_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "emb_shift_invert_free": (C.c_int, [C.c_void_p]),
    "emb_topology_build": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                     C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "emb_topology_get": (C.c_int, [C.c_void_p] + [C.c_void_p] * 10),
    "emb_symbolic": (C.c_int, [C.c_void_p]),
    "emb_assemble_KM": (C.c_int, [C.c_void_p]),
    "emb_assemble_config": (C.c_int, [C.c_void_p, C.c_int64, C.c_int]),
    "emb_assemble_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "emb_n_field": (C.c_int64, [C.c_void_p]),
    "emb_nnz": (C.c_int64, [C.c_void_p]),
    "emb_get_csr": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "emb_csr_rows": (C.c_int64, [C.c_void_p, C.c_int]),
    "emb_csr_nnz": (C.c_int64, [C.c_void_p, C.c_int]),
    "emb_element_matrices": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "emb_surface_define": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "emb_surface_points": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "emb_surface_set_U": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "emb_surface_blocks": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "emb_set_dirichlet": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p]),
    "emb_n_solve": (C.c_int64, [C.c_void_p]),
    "emb_get_solve_ids": (C.c_int, [C.c_void_p, C.c_void_p]),
    "emb_get_solve_perm": (C.c_int, [C.c_void_p, C.c_void_p]),
    "emb_is_paired": (C.c_int, [C.c_void_p]),
    "emb_form_A": (C.c_int, [C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p]),
    "emb_spmv_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "emb_spmv_bench": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double)]),
    "emb_aux_clear": (C.c_int, [C.c_void_p]),
    "emb_aux_add": (C.c_int, [C.c_void_p, C.c_int64] + [C.c_void_p] * 6),
    "emb_aux_add_ex": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64] + [C.c_void_p] * 6 + [C.c_int] * 4),
    "emb_aux_build_top": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p]),
    "emb_aux_get": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_void_p,
                              C.c_void_p, C.c_void_p]),
    "emb_amg_create": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "emb_amg_add_level": (C.c_int, [C.c_void_p, C.c_int, C.c_int64] + [C.c_void_p] * 4 + [C.c_double, C.c_int64] + [C.c_void_p] * 6),
    "emb_amg_set_coarse_inverse": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p]),
    "emb_spmv_sampled": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "emb_solve": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(SolveOpts), C.c_void_p, C.POINTER(SolveInfo)]),
    "emb_solve_multi": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(SolveOpts), C.c_void_p, C.c_void_p]),
    "emb_select_solution": (C.c_int, [C.c_void_p, C.c_int]),
    "emb_fields_async": (C.c_int, [C.c_void_p, C.c_int]),
    "emb_fields_sync": (C.c_int, [C.c_void_p]),
    "emb_precond_sampled": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "emb_spmv_bench_ex": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "emb_solver_config": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "emb_graph_launch_count": (C.c_int64, [C.c_void_p]),
    "emb_solver_block": (C.c_int, [C.c_void_p, C.c_int]),
    "emb_solver_coarse_basis": (C.c_int, [C.c_void_p, C.c_int]),
    "emb_solve_rhs": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(SolveOpts), C.c_void_p, C.POINTER(SolveInfo)]),
    "emb_recycle_config": (C.c_int, [C.c_void_p, C.c_int, C.c_double]),
    "emb_recycle_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_double)]),
    "emb_recycle_accepted": (C.c_int64, [C.c_void_p]),
    "emb_recycle_export": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "emb_recycle_import": (C.c_int, [C.c_void_p, C.c_void_p]),
    "emb_interp": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "emb_interp_last": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lib = None


class EmergeB200Error(RuntimeError):
    """Raised on a negative status of the C ABI (the wrapper's analogue of fem SimulationError)."""


class NotConverged(EmergeB200Error):
    pass


def exported_symbols():
    return sorted(_SIGS)


def load_library():
    """dlopen the CUDA library and bind every symbol of the header; raises if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EmergeB200Error(f"{LIB_PATH} not found: run `python -m emerge_b200.build` (nvcc, sm_100a). "
                              "There is no CPU fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if os.environ.get("EMB_API_TRACE"):
        lib = _TracedLib(lib)
    _lib = lib
    return lib


API_TRACE = {}      # symbol -> [calls, host wall seconds], filled when EMB_API_TRACE is set (diagnostic of the end-to-end legs)


class _TracedLib:
    """Proxy of the loaded library that accumulates the host wall time spent inside every entry point."""

    def __init__(self, lib):
        import time
        for name in _SIGS:
            fn = getattr(lib, name)

            def timed(*a, _fn=fn, _name=name, _clock=time.perf_counter):
                t0 = _clock()
                try:
                    return _fn(*a)
                finally:
                    rec = API_TRACE.setdefault(_name, [0, 0.0])
                    rec[0] += 1
                    rec[1] += _clock() - t0
            setattr(self, name, timed)


def api_trace_report(reset=True, top=14):
    rows = sorted(API_TRACE.items(), key=lambda kv: -kv[1][1])[:top]
    out = ", ".join(f"{k}:{v[0]}x{v[1]:.3f}s" for k, v in rows)
    if reset:
        API_TRACE.clear()
    return out


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class Context:
    """One GPU context (opaque handle of the C ABI)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.emb_create(device, C.byref(h))
        if rc != 0 or not h:
            raise EmergeB200Error(f"emb_create(device={device}) failed (rc={rc}): no usable CUDA device; "
                                  "this path has no CPU fallback")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.emb_destroy(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc, allow_positive=False):
        if rc < 0:
            raise EmergeB200Error(self.lib.emb_last_error(self.h).decode())
        if rc > 0 and not allow_positive:
            raise NotConverged(self.lib.emb_last_error(self.h).decode())
        return rc

    # ---- info
    @property
    def launches(self) -> int:
        return int(self.lib.emb_launch_count(self.h))

    def last_ms(self, phase: str) -> float:
        return float(self.lib.emb_last_ms(self.h, phase.encode()))

    def timer_start(self):
        self._check(self.lib.emb_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_double()
        self._check(self.lib.emb_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def profiler(self, on: bool):
        """cudaProfilerStart/Stop window for `ncu --profile-from-start off`."""
        self._check(self.lib.emb_profiler(self.h, int(bool(on))))

    @property
    def n_field(self) -> int:
        return int(self.lib.emb_n_field(self.h))

    @property
    def nnz(self) -> int:
        return int(self.lib.emb_nnz(self.h))

    @property
    def n_solve(self) -> int:
        return int(self.lib.emb_n_solve(self.h))

    # ---- mesh / assembly
    def upload_mesh(self, nodes, tets, tris, tet_to_field, tri_to_field, n_edges):
        """nodes (3,nN); tets (4,nT); tris (3,nTri); tet_to_field (20,nT); tri_to_field (8,nTri) as the reference
        holds them (F-order views are passed without a copy when possible)."""
        nodes_n3 = _c(np.asarray(nodes).T, np.float64)
        tets_n4 = _c(np.asarray(tets).T, np.int64)
        tris_n3 = _c(np.asarray(tris).T, np.int64)
        ttf = _c(tet_to_field, np.int64)
        t2f = _c(tri_to_field, np.int64)
        self._check(self.lib.emb_upload_mesh(self.h, nodes_n3.shape[0], tets_n4.shape[0], int(n_edges), tris_n3.shape[0],
                                             _p(nodes_n3), _p(tets_n4), _p(tris_n3), _p(ttf), _p(t2f)))
        self.n_tets = tets_n4.shape[0]

    def mesh_tables(self, nodes_xyz, tets_n4, edges=None, tris=None):
        """Device counterpart of Mesh3D.update() + Nedelec2.__init__ (fem/mesh3d.py:224-355, fem/elements/nedelec2.py:32-62):
        -> emerge_b200.synthmesh.MeshTables.  edges (2,nE) / tris (3,nTri): the caller's numbering (the reference's), kept
        verbatim; None: lexicographic numbering (what synthmesh.mesh_tables produces on the host)."""
        from .synthmesh import MeshTables
        nodes = np.ascontiguousarray(nodes_xyz, dtype=np.float64)
        tets = np.ascontiguousarray(tets_n4, dtype=np.int64)
        nN, nT = nodes.shape[0], tets.shape[0]
        ge = np.ascontiguousarray(edges, dtype=np.int64) if edges is not None else None
        gt = np.ascontiguousarray(tris, dtype=np.int64) if tris is not None else None
        nE, nTri = C.c_int64(), C.c_int64()
        self._check(self.lib.emb_topology_build(self.h, nN, nT, _p(nodes), _p(tets), ge.shape[1] if ge is not None else 0, _p(ge),
                                                gt.shape[1] if gt is not None else 0, _p(gt), C.byref(nE), C.byref(nTri)))
        nE, nTri = nE.value, nTri.value
        i8 = np.int64
        out = dict(edges=np.empty((2, nE), i8), tris=np.empty((3, nTri), i8), tet_to_edge=np.empty((6, nT), i8),
                   tet_to_tri=np.empty((4, nT), i8), tri_to_edge=np.empty((3, nTri), i8), tri_to_tet=np.empty((2, nTri), i8),
                   edge_lengths=np.empty(nE, np.float64), tet_to_field=np.empty((20, nT), i8),
                   tri_to_field=np.empty((8, nTri), i8), edge_to_field=np.empty((2, nE), i8))
        self._check(self.lib.emb_topology_get(self.h, *[_p(out[k]) for k in (
            "edges", "tris", "tet_to_edge", "tet_to_tri", "tri_to_edge", "tri_to_tet", "edge_lengths", "tet_to_field",
            "tri_to_field", "edge_to_field")]))
        return MeshTables(nodes.T, tets.T, out["edges"], out["tris"], out["tet_to_edge"], out["tet_to_tri"], out["tri_to_edge"],
                          out["tri_to_tet"], out["edge_lengths"], out["tet_to_field"], out["tri_to_field"], out["edge_to_field"])

    def locate(self, pts):
        """tetrahedron of every point (3,n) as the reference's interpolation picks it (last containing tet), -1 outside"""
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        out = np.empty(pts.shape[1], dtype=np.int64)
        self._check(self.lib.emb_locate_points(self.h, pts.shape[1], _p(pts), _p(out)))
        return out

    def interp_fields(self, x_full, pts, tet_ids=None, curl_const=None):
        """-> (E (3,n), H (3,n) or None): EMDataSet.interpolate on the device (emdata.py:181-199).  x_full None: the
        device-resident solution of the last solve; curl_const (nT,) complex: H = curl(E) * curl_const[tet]."""
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        n = pts.shape[1]
        x = None if x_full is None else np.ascontiguousarray(x_full, dtype=np.complex128)
        tid = None if tet_ids is None else np.ascontiguousarray(tet_ids, dtype=np.int64)
        cc = None if curl_const is None else np.ascontiguousarray(curl_const, dtype=np.complex128)
        E = np.empty((3, n), dtype=np.complex128)
        H = np.empty((3, n), dtype=np.complex128) if cc is not None else None
        self._check(self.lib.emb_interp_fields(self.h, _p(x), n, _p(pts), _p(tid), _p(cc), _p(E), _p(H)))
        return E, H

    def stratton_chu(self, E, H, pos, wnormal, theta, phi, k0):
        """far field (E (3,nout), H (3,nout)) of surface samples: stratton_chu_ff (fem/physics/edm/sc.py:27-142)"""
        E, H = _c(E, np.complex128), _c(H, np.complex128)
        pos, wn = _c(pos, np.float64), _c(wnormal, np.float64)
        th, ph = _c(np.atleast_1d(theta), np.float64), _c(np.atleast_1d(phi), np.float64)
        n, nout = E.shape[1], th.shape[0]
        if H.shape != E.shape or pos.shape != (3, n) or wn.shape != (3, n) or ph.shape[0] != nout:
            raise ValueError("stratton_chu: E, H, pos, wnormal must be (3,n); theta, phi of equal length")
        Eo = np.empty((3, nout), dtype=np.complex128)
        Ho = np.empty((3, nout), dtype=np.complex128)
        self._check(self.lib.emb_stratton_chu(self.h, n, _p(E), _p(H), _p(pos), _p(wn), nout, _p(th), _p(ph), float(k0), _p(Eo), _p(Ho)))
        return Eo, Ho

    def bma_element_matrices(self, xy, tris, edges, tri_to_edge, er, ur, k0):
        """(A, B) (nt,14,14) element matrices of the port eigenproblem (nedeleclegrange2.py:223-417) on the device"""
        xy = _c(xy, np.float64)
        tris, edges, t2e = _c(tris, np.int64), _c(edges, np.int64), _c(tri_to_edge, np.int64)
        er, ur = _c(er, np.complex128), _c(ur, np.complex128)
        nt = tris.shape[1]
        if xy.shape[0] != 2 or tris.shape[0] != 3 or edges.shape[0] != 2 or t2e.shape != (3, nt) or er.shape != (3, 3, nt) \
                or ur.shape != (3, 3, nt):
            raise ValueError("bma_element_matrices: xy (2,n), tris (3,nt), edges (2,ne), tri_to_edge (3,nt), er/ur (3,3,nt)")
        A = np.empty((nt, 14, 14), dtype=np.complex128)
        B = np.empty((nt, 14, 14), dtype=np.complex128)
        self._check(self.lib.emb_bma_element_matrices(self.h, nt, xy.shape[1], edges.shape[1], _p(xy), _p(tris), _p(edges), _p(t2e),
                                                      _p(er), _p(ur), float(k0), _p(A), _p(B)))
        return A, B

    def shift_invert_setup(self, A, B, sigma):
        """device-resident operator v -> (A - sigma B)^-1 B v for dense (n,n) A, B (fem/solver.py:311-357)"""
        A, B = _c(A, np.complex128), _c(B, np.complex128)
        n = A.shape[0]
        if A.shape != (n, n) or B.shape != (n, n):
            raise ValueError("shift_invert_setup: A and B must be square and of equal size")
        sigma = complex(sigma)
        self._check(self.lib.emb_shift_invert_setup(self.h, n, _p(A), _p(B), sigma.real, sigma.imag))
        return n

    def shift_invert_apply(self, x):
        x = _c(x, np.complex128).ravel()
        y = np.empty_like(x)
        self._check(self.lib.emb_shift_invert_apply(self.h, _p(x), _p(y)))
        return y

    def shift_invert_free(self):
        self._check(self.lib.emb_shift_invert_free(self.h))

    def upload_materials(self, er, ur):
        er = _c(er, np.complex128)
        ur = _c(ur, np.complex128)
        assert er.shape == (3, 3, self.n_tets) and ur.shape == er.shape
        self._check(self.lib.emb_upload_materials(self.h, _p(er), _p(ur)))

    def symbolic(self):
        self._check(self.lib.emb_symbolic(self.h))

    def assemble_KM(self):
        self._check(self.lib.emb_assemble_KM(self.h))

    def assemble_mode(self, mode: str = "fused"):
        """numeric-phase algorithm: "fused" (default: per-tet records + one warp per edge / face, rows written once) or
        "coo" (element kernel -> COO scratch -> deterministic row reduction)."""
        self._check(self.lib.emb_assemble_mode(self.h, {"fused": 0, "coo": 1}[mode]))

    def assemble_config(self, chunk_tets: int = 0, persist_l2: bool = True):
        """tets per chunk of the numeric phase (0: default = single pass) and L2 pinning of a chunk's COO scratch."""
        self._check(self.lib.emb_assemble_config(self.h, int(chunk_tets), int(bool(persist_l2))))

    def get_csr(self, which: int, values: bool = True, pattern: bool = True):
        rows = int(self.lib.emb_csr_rows(self.h, which))
        nnz = int(self.lib.emb_csr_nnz(self.h, which))
        indptr = np.empty(rows + 1, dtype=np.int64) if pattern else None
        indices = np.empty(nnz, dtype=np.int32) if pattern else None
        data = np.empty(nnz, dtype=np.complex128) if values else None
        if which != 2 or not self.paired:
            self._check(self.lib.emb_get_csr(self.h, which, _p(indptr), _p(indices), _p(data)))
            return indptr, indices, data
        # the library holds A(f) in solve-index (pair) order; callers get it in the reference's ascending-dof order
        import scipy.sparse as sp
        ip = np.empty(rows + 1, dtype=np.int64)
        ix = np.empty(nnz, dtype=np.int32)
        self._check(self.lib.emb_get_csr(self.h, which, _p(ip), _p(ix), _p(data)))
        perm = self.solve_perm()
        vals = data if data is not None else np.ones(nnz, dtype=np.complex128)
        A = sp.csr_matrix((vals, ix, ip), shape=(rows, rows))[perm][:, perm].tocsr()
        A.sort_indices()
        return (A.indptr.astype(np.int64) if pattern else None, A.indices.astype(np.int32) if pattern else None,
                A.data if values else None)

    def element_matrices(self, t0, t1):
        E = np.empty((t1 - t0, 20, 20), dtype=np.complex128)
        B = np.empty_like(E)
        self._check(self.lib.emb_element_matrices(self.h, t0, t1, _p(E), _p(B)))
        return E, B

    # ---- surfaces
    def surface_define(self, sid, tri_ids, frame=0, basis_inv=None, origin=None):
        tri_ids = _c(tri_ids, np.int64)
        bi = _c(basis_inv if basis_inv is not None else np.eye(3), np.float64)
        og = _c(origin if origin is not None else np.zeros(3), np.float64)
        self._check(self.lib.emb_surface_define(self.h, sid, len(tri_ids), _p(tri_ids), frame, _p(bi), _p(og)))
        return len(tri_ids)

    def surface_points(self, sid, ntri):
        xy = np.empty((2, 6, ntri), dtype=np.float64)
        self._check(self.lib.emb_surface_points(self.h, sid, _p(xy)))
        return xy

    def surface_set_U(self, sid, U, want_full=False):
        U = _c(U, np.complex128)
        b = np.empty(self.n_field, dtype=np.complex128) if want_full else None
        self._check(self.lib.emb_surface_set_U(self.h, sid, _p(U), _p(b)))
        return b

    def surface_blocks(self, sid, ntri):
        S = np.empty((ntri, 8, 8), dtype=np.float64)
        self._check(self.lib.emb_surface_blocks(self.h, sid, _p(S)))
        return S

    # ---- elimination / operator
    def set_dirichlet(self, pec_ids):
        ids = _c(pec_ids, np.int64)
        self._check(self.lib.emb_set_dirichlet(self.h, len(ids), _p(ids)))
        self._perm = None

    @property
    def paired(self) -> bool:
        """True when the library numbers the solve space by (edge / face) function pairs (block-CSR operator)."""
        return bool(self.lib.emb_is_paired(self.h))

    def solve_perm(self):
        """perm[s] = internal solve index of solve_ids()[s] (identity unless `paired`)."""
        if getattr(self, "_perm", None) is None:
            out = np.empty(self.n_solve, dtype=np.int64)
            self._check(self.lib.emb_get_solve_perm(self.h, _p(out)))
            self._perm = out
        return self._perm

    def _rows_to_internal(self, R):
        """rows of a solve-space matrix given in ascending-dof order -> solve-index order"""
        perm = self.solve_perm()
        inv = np.empty_like(perm)
        inv[perm] = np.arange(len(perm))
        return R[inv]

    def solve_ids(self):
        out = np.empty(self.n_solve, dtype=np.int64)
        self._check(self.lib.emb_get_solve_ids(self.h, _p(out)))
        return out

    def form_A(self, k0, sids=(), gammas=()):
        s = _c(list(sids), np.int32)
        g = _c(list(gammas), np.complex128)
        self._check(self.lib.emb_form_A(self.h, float(k0), len(s), _p(s), _p(g)))

    def spmv(self, x):
        """y = A(f) x on the solve space, x and y in the order of solve_ids()"""
        perm = self.solve_perm()
        xi = np.empty(self.n_solve, dtype=np.complex128)
        xi[perm] = _c(x, np.complex128)
        y = np.empty_like(xi)
        self._check(self.lib.emb_spmv_host(self.h, _p(xi), _p(y)))
        return y[perm]

    def spmv_bench(self, reps=20, nv=1, fp32=False):
        """avg ms of one operator application on resident vectors (nv interleaved right-hand sides; fp32: the complex64
        inner operator instead of A(f))"""
        ms = C.c_double()
        self._check(self.lib.emb_spmv_bench_ex(self.h, reps, int(nv), int(bool(fp32)), C.byref(ms)))
        return ms.value

    def solver_config(self, inner_fp32=True, side_streams=True, block=True):
        """inner_fp32: complex64 storage of the inner operator; side_streams: concurrent auxiliary spaces;
        block: the ports of a lockstep group share one Krylov space (block COCR)"""
        self._check(self.lib.emb_solver_config(self.h, int(bool(inner_fp32)), int(bool(side_streams))))
        self._check(self.lib.emb_solver_block(self.h, int(bool(block))))

    def solver_coarse_basis(self, on: bool):
        """EXPERIMENTAL (default off): the reduced basis as an extra coarse space of the preconditioner; clears the basis."""
        self._check(self.lib.emb_solver_coarse_basis(self.h, int(bool(on))))

    @property
    def graph_launches(self) -> int:
        return int(self.lib.emb_graph_launch_count(self.h))

    def precond_sampled(self):
        ms, cnt = C.c_double(), C.c_int64()
        self._check(self.lib.emb_precond_sampled(self.h, C.byref(ms), C.byref(cnt)))
        return ms.value, cnt.value

    def aux_clear(self):
        self._check(self.lib.emb_aux_clear(self.h))
        self._n_aux = 0

    @staticmethod
    def _csr_args(M):
        M = M.tocsr()
        if M.dtype != np.float64:
            M = M.astype(np.float64)
        if not M.has_sorted_indices:
            M.sort_indices()
        return [_c(M.indptr, np.int64), _c(M.indices, np.int32), _c(M.data, np.float64)]

    @staticmethod
    def csr_pair(R):
        """(arguments of R, arguments of R^T) as the aux_add* calls take them; pure host work, safe to run in a worker
        thread while the owner thread talks to the device"""
        return Context._csr_args(R) + Context._csr_args(R.T)

    def aux_add(self, R, rows_internal=False, prepared=None):
        """R: scipy sparse (n_solve x ncol) real transfer matrix of one auxiliary space, rows in the order of solve_ids()
        (or already in the library's solve-index order when rows_internal).  prepared: csr_pair(R) computed elsewhere."""
        return self.aux_add_ex(R, parent=-1, solver="diag", rows_internal=rows_internal, prepared=prepared)

    def aux_add_ex(self, R, parent=-1, solver="diag", hid=-1, scale="one", rows_internal=False, prepared=None):
        """R: scipy sparse (rows of the parent space x ncol) real transfer matrix; returns the index of the new space.
        Top-level spaces (parent < 0): rows in the order of solve_ids(), or in solve-index order when rows_internal."""
        if prepared is None:
            if parent < 0 and not rows_internal:
                R = self._rows_to_internal(R.tocsr())
            prepared = self.csr_pair(R)
        a = prepared
        self._check(self.lib.emb_aux_add_ex(self.h, R.shape[0], R.shape[1], *[_p(x) for x in a], int(parent),
                                            {"diag": 0, "amg": 1}[solver], int(hid), {"one": 0, "minus_inv_k0sq": 1}[scale]))
        self._n_aux = getattr(self, "_n_aux", 0) + 1
        return self._n_aux - 1

    def aux_build_top(self, which: str, edges, n_bad: int):
        """builds the top-level space 'G' (P2 gradients) or 'P' (Whitney) on the device and appends it (emb_aux_build_top).
        Returns (index of the new space or -1 when no column is left, ncol, bad (n_bad,) bool: dropped columns)."""
        from .auxspace import _face_tables
        vt, et, wt = _face_tables()
        tab = _c(np.concatenate([vt.ravel(), et.ravel(), wt.ravel()]), np.float64)
        edges = _c(edges, np.int64)
        ncol = C.c_int64()
        bad = np.zeros(int(n_bad), dtype=np.uint8)
        self._check(self.lib.emb_aux_build_top(self.h, {"G": 0, "P": 1}[which], _p(edges), _p(tab), C.byref(ncol), _p(bad)))
        if ncol.value == 0:
            return -1, 0, bad.astype(bool)
        self._n_aux = getattr(self, "_n_aux", 0) + 1
        return self._n_aux - 1, int(ncol.value), bad.astype(bool)

    def aux_get(self, idx: int):
        """scipy CSR of auxiliary space idx's transfer matrix, rows in solve-index order"""
        import scipy.sparse as sp
        nrow, ncol, nnz = C.c_int64(), C.c_int64(), C.c_int64()
        self._check(self.lib.emb_aux_get(self.h, int(idx), C.byref(nrow), C.byref(ncol), C.byref(nnz), None, None, None))
        ip = np.empty(nrow.value + 1, dtype=np.int64)
        ix = np.empty(nnz.value, dtype=np.int32)
        dv = np.empty(nnz.value, dtype=np.float64)
        self._check(self.lib.emb_aux_get(self.h, int(idx), C.byref(nrow), C.byref(ncol), C.byref(nnz), _p(ip), _p(ix), _p(dv)))
        return sp.csr_matrix((dv, ix, ip), shape=(nrow.value, ncol.value))

    def amg_upload(self, levels):
        """levels: output of emerge_b200.amg.sa_hierarchy (finest first); returns the hierarchy id."""
        hid = C.c_int()
        self._check(self.lib.emb_amg_create(self.h, C.byref(hid)))
        for lev in levels:
            A, P = lev["A"], lev["P"]
            if P is None:
                n = A.shape[0]
                self._check(self.lib.emb_amg_add_level(self.h, hid.value, n, None, None, None, None, 1.0, 0, *([None] * 6)))
                inv = _c(np.linalg.inv(A.toarray()), np.float64)
                self._check(self.lib.emb_amg_set_coarse_inverse(self.h, hid.value, n, _p(inv)))
                break
            a = self._csr_args(A) + [_c(lev["dinv"], np.float64)]
            pt = self._csr_args(P) + self._csr_args(P.T)
            self._check(self.lib.emb_amg_add_level(self.h, hid.value, A.shape[0], *[_p(x) for x in a],
                                                   float(4.0 / (3.0 * lev["rho"])), P.shape[1], *[_p(x) for x in pt]))
        return hid.value

    def spmv_sampled(self):
        ms, cnt = C.c_double(), C.c_int64()
        self._check(self.lib.emb_spmv_sampled(self.h, C.byref(ms), C.byref(cnt)))
        return ms.value, cnt.value

    # ---- subspace recycling across frequency points
    def recycle_config(self, max_vectors: int, snapshot_rtol_factor: float = 0.3):
        self._check(self.lib.emb_recycle_config(self.h, int(max_vectors), float(snapshot_rtol_factor)))

    def recycle_info(self):
        n, sp, rr = C.c_int(), C.c_int64(), C.c_double()
        self._check(self.lib.emb_recycle_info(self.h, C.byref(n), C.byref(sp), C.byref(rr)))
        return dict(n=n.value, spmvs=sp.value, last_proj_relres=rr.value)

    def recycle_accepted(self) -> int:
        """monotonic count of directions accepted into the reduced basis (survives compaction and resets)"""
        return int(self.lib.emb_recycle_accepted(self.h))

    def recycle_export(self, j: int, device_ptr: int):
        self._check(self.lib.emb_recycle_export(self.h, int(j), C.c_void_p(device_ptr)))

    def recycle_import(self, device_ptr: int):
        self._check(self.lib.emb_recycle_import(self.h, C.c_void_p(device_ptr)))

    @staticmethod
    def _opts(method="cocr", precond="block", restart=50, maxit=100000, rtol=1e-8, use_x0=False):
        return SolveOpts(METHODS[method], PRECONDS[precond], restart, maxit, rtol, int(use_x0))

    def solve(self, sid, want_x=True, raise_on_fail=True, out=None, **kw):
        o = self._opts(**kw)
        info = SolveInfo()
        x = out if out is not None else (np.zeros(self.n_field, dtype=np.complex128) if want_x else None)
        rc = self._check(self.lib.emb_solve(self.h, sid, C.byref(o), _p(x), C.byref(info)), allow_positive=not raise_on_fail)
        return x, dict(iters=info.iters, relres=info.relres, ms=info.ms, spmvs=info.spmvs, converged=rc == 0)

    def solve_multi(self, sids, want_x=True, raise_on_fail=True, outs=None, **kw):
        """All right-hand sides `sids` (1..4 surfaces) of the current A(f) in lockstep.  Returns (list of x or None, list
        of info dicts).  outs: optional list of preallocated complex128[n_field] host buffers (None entries allowed)."""
        sids = [int(s) for s in sids]
        n = len(sids)
        o = self._opts(**kw)
        infos = (SolveInfo * n)()
        xs = []
        for k in range(n):
            x = outs[k] if outs is not None and outs[k] is not None else (np.zeros(self.n_field, dtype=np.complex128) if want_x else None)
            xs.append(x)
        ptrs = (C.c_void_p * n)(*[None if x is None else x.ctypes.data for x in xs])
        sid_arr = (C.c_int * n)(*sids)
        rc = self._check(self.lib.emb_solve_multi(self.h, n, sid_arr, C.byref(o), ptrs, infos), allow_positive=not raise_on_fail)
        out = [dict(iters=i.iters, relres=i.relres, ms=i.ms, spmvs=i.spmvs, converged=bool(i.relres <= o.rtol)) for i in infos]
        if rc == 0:
            for d in out:
                d["converged"] = True
        return xs, out

    def fields_async(self, on: bool = True):
        """host copies of the solved fields (solve_multi(outs=...)) overlap the following work; valid after fields_sync()"""
        self._check(self.lib.emb_fields_async(self.h, 1 if on else 0))

    def fields_sync(self):
        self._check(self.lib.emb_fields_sync(self.h))

    def select_solution(self, k: int):
        """column k of the last solve_multi becomes the device-resident solution used by interp(None, ...)"""
        self._check(self.lib.emb_select_solution(self.h, int(k)))

    def solve_rhs(self, b_full, x0=None, raise_on_fail=True, **kw):
        b = _c(b_full, np.complex128)
        o = self._opts(use_x0=x0 is not None, **kw)
        info = SolveInfo()
        x = np.zeros(self.n_field, dtype=np.complex128) if x0 is None else _c(x0, np.complex128).copy()
        rc = self._check(self.lib.emb_solve_rhs(self.h, _p(b), C.byref(o), _p(x), C.byref(info)), allow_positive=not raise_on_fail)
        return x, dict(iters=info.iters, relres=info.relres, ms=info.ms, spmvs=info.spmvs, converged=rc == 0)

    def interp(self, x_full, tet_ids, xyz):
        """E (3,npts) at points xyz (3,npts), point k inside tet tet_ids[k]; x_full=None uses the device solution."""
        tet_ids = _c(tet_ids, np.int64)
        xyz = _c(xyz, np.float64)
        E = np.empty((3, len(tet_ids)), dtype=np.complex128)
        if x_full is None:
            self._check(self.lib.emb_interp_last(self.h, len(tet_ids), _p(tet_ids), _p(xyz), _p(E)))
        else:
            x = _c(x_full, np.complex128)
            self._check(self.lib.emb_interp(self.h, _p(x), len(tet_ids), _p(tet_ids), _p(xyz), _p(E)))
        return E
