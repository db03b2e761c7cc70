"""Frequency sweep driver: host mirror of Electrodynamics3D.frequency_domain (fem/physics/edm/emfreq3d.py:607-732)
over the CUDA library.  Everything frequency-independent (K, M, the surface matrices S_p, the eliminated
pattern, the S-parameter sample points) is set up once; per frequency only gamma_p(f), the incident-field
samples U_p(f) and the solves cross the boundary.

S-parameter extraction mirrors emfreq3d.py:734-779, fem/mth/sparam.py:72-139 and fem/mth/integrals.py:24-70:
the field is evaluated on the GPU at the port's Dunavant-4 points (emb_interp) in the tetrahedron the reference's
interpolation would pick (fem/mth/tet.py:393-497: the LAST tetrahedron of the candidate list containing the
point wins), the small surface sums are done on the host exactly as the reference does them.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field

import numpy as np

from .lib import Context

C0 = 299792458.0

# gaus_quad_tri(4): rows W, L1, L2, L3 (fem/mth/optimized.py:28-29,77-108)
_W = (0.223381589678011, 0.109951743655322)
_L = ((0.108103018168070, 0.445948490915965, 0.445948490915965),
      (0.816847572980459, 0.091576213509771, 0.091576213509771))


def dunavant4() -> np.ndarray:
    rows = []
    for w, (l1, l2, l3) in zip(_W, _L):
        for _ in range(3):
            rows.append([w, l1, l2, l3])
            l1, l2, l3 = l2, l3, l1
    return np.array(rows).T


def hierarchical_order(n: int) -> list:
    """Processing order of n frequency points for a sweep with subspace recycling: both ends first, then midpoints of
    ever finer bisection.  Later points are then INTERPOLATED by the recycled space (combination weights O(1)) instead of
    extrapolated (weights growing like binomials, which amplifies the 1e-9 solver noise of the stored directions above
    rtol)."""
    if n <= 2:
        return list(range(n))
    out, seg = [0, n - 1], [(0, n - 1)]
    while seg:
        nxt = []
        for a, b in seg:
            if b - a > 1:
                m = (a + b) // 2
                out.append(m)
                nxt += [(a, m), (m, b)]
        seg = nxt
    return out


def _is(bc, name):
    """isinstance by class name through the MRO so reference objects (fem.bc.*) are accepted too."""
    return any(k.__name__ == name for k in type(bc).__mro__)


def _tri_ids(bc, get_triangles):
    if hasattr(bc, "tri_ids"):
        return np.asarray(bc.tri_ids, dtype=np.int64)
    return np.asarray(get_triangles(bc.tags), dtype=np.int64)


@dataclass
class SweepResult:
    freqs: np.ndarray
    port_numbers: list
    S: np.ndarray                                   # (nf, P, P): S[f, i, j] = port i response, port j excited
    stats: list = field(default_factory=list)       # per (freq, port) solver info dicts
    fields: dict = field(default_factory=dict)      # optional {(ifreq, port_number): x (n_field)}
    timings: dict = field(default_factory=dict)


class FrequencySweep:
    """tables: emerge_b200.synthmesh.MeshTables-like object (or any object exposing the reference's arrays:
    nodes, tets, tris, edges, tri_to_tet, tet_to_field, tri_to_field).  bcs: PEC / RobinBC objects (ours or fem's)."""

    def __init__(self, tables, er, ur, bcs, device: int = 0, get_triangles=None, ctx: Context | None = None,
                 recycle: int = 40, multilevel: bool = True, f_ref: float = 10e9, recycle_snap: float = 0.3,
                 coarse_basis: bool = True):
        self.t = tables
        self.er = np.ascontiguousarray(er, dtype=np.complex128)
        self.ur = np.ascontiguousarray(ur, dtype=np.complex128)
        self.bcs = list(bcs)
        self.get_triangles = get_triangles
        self.ctx = ctx if ctx is not None else Context(device)
        self.timings = {}
        self._setup_done = False
        self.multilevel = bool(multilevel)  # AMG V-cycles on the nodal auxiliary problems (False: Jacobi on every space)
        self.f_ref = float(f_ref)           # frequency whose k0^2 shifts the nodal Helmholtz-type auxiliary operator
        self.recycle = int(recycle)        # directions kept from previous frequency points (0 = every point solved cold)
        self.recycle_snap = float(recycle_snap)   # points that iterate are solved to recycle_snap * rtol (they feed the basis)
        import os
        # reduced basis as an extra coarse space of the preconditioner (1M tets, first 20 points of the sweep: 5,130 -> 3,420
        # block iterations, profiles/r2_coarse_basis_1M.json); EMB_COARSE_BASIS=0 switches it off
        self.coarse_basis = bool(coarse_basis) and os.environ.get("EMB_COARSE_BASIS", "1") != "0"
        self.lockstep = 4                   # ports solved together per lockstep group (1 = one port at a time)
        self.amg_coarse_size = 2500         # the AMG level at or below this size is inverted densely (one launch, L2-resident)
        self.solver_opts = dict(method="cocr", precond="multilevel", rtol=1e-8, maxit=200000, restart=50)

    # ------------------------------------------------------------------ setup (once)
    def setup(self):
        t, ctx = self.t, self.ctx
        n_edges = t.edges.shape[1]
        t0 = time.perf_counter()
        # the host pieces of the multilevel setup that only need the mesh tables start now, in worker threads (numpy /
        # scipy release the GIL), and overlap the upload, the symbolic phase and the assembly on the device
        self._early = self._start_early_aux() if self.solver_opts.get("precond") == "multilevel" else None
        ctx.upload_mesh(t.nodes, t.tets, t.tris, t.tet_to_field, t.tri_to_field, n_edges)
        ctx.upload_materials(self.er, self.ur)
        self.timings["upload_s"] = time.perf_counter() - t0
        ctx.symbolic()
        self.timings["symbolic_ms"] = ctx.last_ms("symbolic")
        ctx.assemble_KM()
        self.timings["tet_kernel_ms"] = ctx.last_ms("tet_kernel")
        self.timings["reduce_ms"] = ctx.last_ms("reduce")
        # PEC dofs: all 8 functions of every PEC triangle (its 3 edges x2 + face x2), assembler.py:348-359
        pec = [np.asarray(t.tri_to_field)[:, _tri_ids(b, self.get_triangles)].ravel() for b in self.bcs if _is(b, "PEC")]
        pec = np.unique(np.concatenate(pec)) if pec else np.zeros(0, dtype=np.int64)
        self.robin = [b for b in self.bcs if _is(b, "RobinBC")]
        self.ports = [b for b in self.bcs if _is(b, "PortBC")]
        self.sid = {}
        self.ntri = {}
        self.points = {}
        for sid, b in enumerate(self.robin):
            ids = _tri_ids(b, self.get_triangles)
            if b._include_force:
                n = ctx.surface_define(sid, ids, 0, np.asarray(b.get_inv_basis(), dtype=float), np.asarray(b.cs.origin, dtype=float))
                self.points[id(b)] = ctx.surface_points(sid, n)
            else:
                n = ctx.surface_define(sid, ids, 1)
            self.sid[id(b)] = sid
            self.ntri[id(b)] = n
        ctx.set_dirichlet(pec)
        self.pec_ids = pec
        ctx.recycle_config(self.recycle, self.recycle_snap)
        if self.coarse_basis and self.recycle:
            ctx.solver_coarse_basis(True)
        if self.solver_opts.get("precond") == "multilevel":
            t1 = time.perf_counter()
            self._setup_aux_spaces()
            self.timings["aux_setup_s"] = time.perf_counter() - t1
        self._setup_sample_points()
        self._setup_done = True

    def _setup_sample_points(self):
        """S-parameter sample points per port (host only: generate_int_points_tri / calc_area of the port triangles)"""
        t = self.t
        DP = dunavant4()
        nodes = np.asarray(t.nodes)
        tris = np.asarray(t.tris)
        self._sp = {}
        self._sp_all = None
        for b in self.ports:
            ids = _tri_ids(b, self.get_triangles)
            tv = tris[:, ids]                                           # (3, ntri)
            P = nodes[:, tv]                                            # (3 xyz, 3 vert, ntri)
            pts = np.einsum("kq,xkt->xqt", DP[1:4], P)                 # (3, 6, ntri)  generate_int_points_tri
            e1, e2 = P[:, 1] - P[:, 0], P[:, 2] - P[:, 0]
            area = 0.5 * np.linalg.norm(np.cross(e1.T, e2.T), axis=1)   # calc_area
            tets = np.asarray(t.tri_to_tet)[:, ids].max(axis=0)        # last containing tet wins
            tet0 = np.asarray(t.tri_to_tet)[0, ids]                     # ertri/urtri use tri_to_tet[0] (emfreq3d.py:621-624)
            self._sp[id(b)] = dict(pts=pts.reshape(3, -1), tet=np.repeat(tets[None, :], 6, axis=0).ravel(), area=area,
                                   tet0=tet0, ntri=len(ids))
            if getattr(b, "v_integration", False):
                self._setup_vline(b, ids)

    def _start_early_aux(self):
        from concurrent.futures import ThreadPoolExecutor
        from .auxspace import nodal_interpolation, p1_gradient, p1_stiffness_mass
        if not self.multilevel:
            return None
        t = self.t
        tr = lambda T: np.real(T[0, 0] + T[1, 1] + T[2, 2]) / 3.0
        w_eps, w_mu = tr(self.er), 1.0 / tr(self.ur)
        same = bool(np.allclose(w_eps, w_mu))
        ex = ThreadPoolExecutor(max_workers=4)
        return dict(pool=ex, w_eps=w_eps, w_mu=w_mu, same=same,
                    f_le=ex.submit(p1_stiffness_mass, t, w_eps),
                    f_lm=None if same else ex.submit(p1_stiffness_mass, t, w_mu),
                    f_pi=ex.submit(nodal_interpolation, t), f_g1=ex.submit(p1_gradient, t))

    def _setup_aux_spaces(self):
        """Auxiliary spaces of the additive multilevel preconditioner, restricted to the solve space; columns whose
        support touches an eliminated dof are dropped (their potential / Whitney / nodal dof is fixed by the PEC condition).

            M^-1 = D_blk^-1 + G D_G^-1 G^T + P [ D_P^-1 + G1 (-k0^-2 V_eps) G1^T + sum_c Pi_c V_mu Pi_c^T ] P^T

        G: gradients of the P2 Lagrange space (the kernel of the curl-curl matrix), P: the Whitney space, G1: gradients
        of P1 inside the Whitney space, Pi_c: nodal vector fields (Hiptmair-Xu); V_eps / V_mu are V-cycles of
        smoothed-aggregation hierarchies of the P1 operators (grad, eps grad) and (grad, mu^-1 grad) + k^2 (lumped mass),
        set up once per mesh (emerge_b200/amg.py).  With multilevel=False the nodal problems are not solved and the
        P1 gradients only get a Jacobi scaling (the round-1 first version)."""
        from concurrent.futures import ThreadPoolExecutor
        import scipy.sparse as sp
        from .auxspace import build_aux_spaces, build_aux_spaces_paired, nodal_interpolation, p1_stiffness_mass
        from .amg import sa_hierarchy
        ctx = self.ctx
        t = self.t
        N = 2 * t.edges.shape[1] + 2 * t.tris.shape[1]
        keep = np.ones(N, dtype=bool)
        keep[self.pec_ids] = False
        multilevel = self.multilevel
        early = getattr(self, "_early", None)
        self._early = None
        if early is not None:
            w_eps, w_mu, same = early["w_eps"], early["w_mu"], early["same"]
        else:
            tr = lambda T: np.real(T[0, 0] + T[1, 1] + T[2, 2]) / 3.0
            w_eps, w_mu = tr(self.er), 1.0 / tr(self.ur)
            same = np.allclose(w_eps, w_mu)
        kmid2 = (2 * np.pi * self.f_ref / C0) ** 2

        paired = ctx.paired                           # device queries from the owner thread only
        perm = None if paired else ctx.solve_perm()

        def top_level():
            if paired:
                # rows straight in the library's pair order of the solve space, no full matrices (auxspace.py)
                return build_aux_spaces_paired(t, keep)
            G, P, G1 = build_aux_spaces(t)
            elim = ~keep
            inv = np.empty_like(perm)
            inv[perm] = np.arange(len(perm))
            rows_int = np.nonzero(keep)[0][inv]

            def restrict(R):
                bad = np.asarray(abs(R[elim]).sum(axis=0)).ravel() > 0
                return R[rows_int][:, ~bad].tocsr(), bad
            Gs, _ = restrict(G)
            Ps, badP = restrict(P)
            return Gs, Ps, badP, G1

        # The host work is numpy / scipy kernels that release the GIL: independent pieces run in worker threads while
        # this (owner) thread is the only one that talks to the device context.
        import os
        # top-level spaces G and P straight on the device (csrc/auxbuild.cu; same entries as build_aux_spaces_paired, which was
        # the critical path of this setup: 3-4 s at 1M tets); EMB_AUX_HOST=1 keeps the numpy builder
        on_device = paired and multilevel and hasattr(ctx, "aux_build_top") and not int(os.environ.get("EMB_AUX_HOST", "0"))
        with (early["pool"] if early is not None else ThreadPoolExecutor(max_workers=4)) as ex:
            f_top = None if on_device else ex.submit(top_level)
            if early is not None:
                f_le, f_lm, f_pi = early["f_le"], early["f_lm"], early["f_pi"]
            else:
                f_le = ex.submit(p1_stiffness_mass, t, w_eps) if multilevel else None
                f_lm = ex.submit(p1_stiffness_mass, t, w_mu) if (multilevel and not same) else None
                f_pi = ex.submit(nodal_interpolation, t) if multilevel else None
            ctx.aux_clear()
            self.aux_dims = []
            if on_device:
                from .auxspace import p1_gradient
                f_g1 = early["f_g1"] if early is not None else ex.submit(p1_gradient, t)
                nN, nE = np.asarray(t.nodes).shape[1], np.asarray(t.edges).shape[1]
                ig, ncolG, _ = ctx.aux_build_top("G", t.edges, nN + nE)
                ip, ncolP, badP = ctx.aux_build_top("P", t.edges, nE)
                if ig >= 0:
                    self.aux_dims.append(ncolG)
                if ip < 0:
                    return
                self.aux_dims.append(ncolP)
                G1 = f_g1.result()
                Gs = Ps = f_g = f_p = None
            else:
                Gs, Ps, badP, G1 = f_top.result()
                f_g = ex.submit(ctx.csr_pair, Gs) if Gs.shape[1] > 0 else None
                f_p = ex.submit(ctx.csr_pair, Ps) if Ps.shape[1] > 0 else None
                ncolP = Ps.shape[1]
            badN = np.asarray(abs(G1[badP]).sum(axis=0)).ravel() > 0
            kn = ~badN
            G1s = G1[~badP][:, kn].tocsr()
            f_he = f_hm = None
            if multilevel and ncolP > 0 and G1s.shape[1] > 0:
                Le, mass = f_le.result()
                Lm = Le if same else f_lm.result()[0]
                f_he = ex.submit(lambda: sa_hierarchy((Le[kn][:, kn] + 1e-3 * kmid2 * sp.diags(mass[kn] * np.mean(w_eps))).tocsr(),
                                                      coarse_size=self.amg_coarse_size))
                f_hm = ex.submit(lambda: sa_hierarchy((Lm[kn][:, kn] + kmid2 * sp.diags(mass[kn] * np.mean(w_mu))).tocsr(),
                                                      coarse_size=self.amg_coarse_size))
                f_kids = [ex.submit(lambda M: ctx.csr_pair(M), G1s)]
                f_pcs = ex.submit(lambda: [Pc[~badP][:, kn].tocsr() for Pc in f_pi.result()])
            if not on_device:
                if f_g is not None:
                    ctx.aux_add(Gs, rows_internal=True, prepared=f_g.result())
                    self.aux_dims.append(Gs.shape[1])
                if f_p is None:
                    return
                if f_he is None:
                    ctx.aux_add(Ps, rows_internal=True, prepared=f_p.result())
                    self.aux_dims.append(Ps.shape[1])
                    if G1s.shape[1] > 0:
                        ctx.aux_add((Ps @ G1s).tocsr(), rows_internal=True)
                        self.aux_dims.append(G1s.shape[1])
                    return
                ip = ctx.aux_add_ex(Ps, parent=-1, solver="diag", rows_internal=True, prepared=f_p.result())
                self.aux_dims.append(Ps.shape[1])
            elif f_he is None:
                return
            He, Hm = f_he.result(), f_hm.result()
            he, hm = ctx.amg_upload(He), ctx.amg_upload(Hm)
            self.amg_levels = dict(eps=[l["A"].shape[0] for l in He], mu=[l["A"].shape[0] for l in Hm])
            ctx.aux_add_ex(G1s, parent=ip, solver="amg", hid=he, scale="minus_inv_k0sq", prepared=f_kids[0].result())
            for Pc in f_pcs.result():
                ctx.aux_add_ex(Pc, parent=ip, solver="amg", hid=hm, scale="one")
            self.aux_dims += [G1s.shape[1]] * 4

    def _setup_vline(self, b, ids):
        """define_lumped_port_integration_points (emfreq3d.py:366-389) + point location for the 10 midpoints."""
        t = self.t
        nodes = np.asarray(t.nodes)
        direction = np.asarray(b.direction.np if hasattr(b.direction, "np") else b.direction, dtype=float)
        pts = np.unique(np.asarray(t.tris)[:, ids])
        dotp = nodes[0, pts] * direction[0] + nodes[1, pts] * direction[1] + nodes[2, pts] * direction[2]
        start = nodes[:, pts[dotp == dotp.min()]].mean(axis=1)
        end = start + direction * b.height
        lin = np.linspace(start, end, 11)                               # Line.from_points(start, end, 11)
        mid = 0.5 * (lin[:-1] + lin[1:])
        d = lin[1:] - lin[:-1]
        # candidate tets = those touching a port vertex (emfreq3d.py:648-653); last containing tet wins
        tets = np.asarray(t.tets)
        allv = np.zeros(nodes.shape[1], dtype=bool)
        for p in self.ports:
            allv[np.unique(np.asarray(t.tris)[:, _tri_ids(p, self.get_triangles)])] = True
        cand = np.nonzero(allv[tets].any(axis=0))[0]
        v0 = nodes[:, tets[0, cand]].T
        Bm = np.stack([nodes[:, tets[k, cand]].T - v0 for k in (1, 2, 3)], axis=2)     # (nc,3,3) columns
        inv = np.linalg.inv(Bm)
        loc = np.einsum("cij,pcj->pci", inv, mid[:, None, :] - v0[None, :, :])          # (10, nc, 3)
        inside = (loc.sum(axis=2) <= 1.00000001) & (loc >= -1e-6).all(axis=2)          # tet.py:425
        tet_of = np.array([cand[np.nonzero(inside[k])[0][-1]] if inside[k].any() else -1 for k in range(mid.shape[0])])
        self._sp[id(b)]["vline"] = dict(mid=mid.T.copy(), d=d, tet=tet_of)

    # ------------------------------------------------------------------ per frequency
    def _port_constants(self, b, sp):
        tet0 = sp["tet0"]
        er, ur = self.er, self.ur
        tr_u = ur[0, 0, tet0] + ur[1, 1, tet0] + ur[2, 2, tet0]
        tr_e = er[0, 0, tet0] + er[1, 1, tet0] + er[2, 2, tet0]
        mt = b.modetype
        if mt == "TEM":
            return 1 / np.sqrt(tr_u / tr_e)
        if mt == "TE":
            return 1 / (tr_u / 3)
        return 1 / (tr_e / 3)

    def _interp_all(self, x_full):
        """E at the S-parameter sample points of EVERY port with one device call (one kernel, one read-back instead of
        one per port pair) -> {id(port): E (3, n_port_points)}; x_full None: the device-resident solution."""
        lay = getattr(self, "_sp_all", None)
        if lay is None:
            tets, pts, sl, o = [], [], {}, 0
            for b in self.ports:
                sp = self._sp[id(b)]
                if getattr(b, "v_integration", False):
                    ok = sp["vline"]["tet"] >= 0
                    tt, pp = sp["vline"]["tet"][ok], sp["vline"]["mid"][:, ok]
                else:
                    tt, pp = sp["tet"], sp["pts"]
                tets.append(np.asarray(tt, dtype=np.int64))
                pts.append(np.asarray(pp, dtype=np.float64))
                sl[id(b)] = (o, o + len(tt))
                o += len(tt)
            lay = self._sp_all = (np.concatenate(tets), np.ascontiguousarray(np.concatenate(pts, axis=1)), sl)
        tets, pts, sl = lay
        E = self.ctx.interp(x_full, tets, pts) if len(tets) else np.zeros((3, 0), dtype=np.complex128)
        return {k: E[:, a:b] for k, (a, b) in sl.items()}

    def _s_data(self, b, k0, x_full, E_all=None, modes=None):
        """(pfield, pmode) of _compute_s_data (emfreq3d.py:734-779).  E_all: the output of _interp_all for this solution;
        modes: a dict shared by the calls of one frequency point (mode field, impedance and material constants of a port
        do not depend on which port is excited)."""
        sp = self._sp[id(b)]
        ctx = self.ctx
        if getattr(b, "v_integration", False):
            vl = sp["vline"]
            ok = vl["tet"] >= 0
            E = np.zeros((3, len(ok)), dtype=np.complex128)
            if ok.any():
                E[:, ok] = E_all[id(b)] if E_all is not None else ctx.interp(x_full, vl["tet"][ok], vl["mid"][:, ok])
            V = np.sum(E[0] * vl["d"][:, 0] + E[1] * vl["d"][:, 1] + E[2] * vl["d"][:, 2])
            a, bb = (b.voltage, V - b.voltage) if b.active else (0, V)
            return np.sqrt(bb ** 2 / (2 * b.Z0)), np.sqrt(a ** 2 / (2 * b.Z0))
        got = modes.get(id(b)) if modes is not None else None
        if got is None:
            const = np.squeeze(self._port_constants(b, sp)).astype(np.complex128)
            mode = np.asarray(b.port_mode_3d_global(sp["pts"][0], sp["pts"][1], sp["pts"][2], k0))
            got = (const, mode, b.Zmode(k0))
            if modes is not None:
                modes[id(b)] = got
        const, mode, Z = got
        E = E_all[id(b)] if E_all is not None else ctx.interp(x_full, sp["tet"], sp["pts"])
        Q = 1 if b.active else 0
        DP = dunavant4()
        f1 = (((E - Q * mode) * np.conj(mode)).sum(axis=0) / (2 * Z)).reshape(6, sp["ntri"])
        f2 = ((mode * np.conj(mode)).sum(axis=0) / (2 * Z)).reshape(6, sp["ntri"])
        pfield = np.sum(const * (DP[0] @ f1) * sp["area"])             # _fast_integral_c
        pmode = np.sum(const * (DP[0] @ f2) * sp["area"])
        return pfield, pmode

    def assemble_frequency(self, freq):
        """Forms A(f) and the port right-hand sides on the device (Assembler.assemble_freq_matrix, assembler.py:312-388)."""
        k0 = 2 * np.pi * freq / 299792458
        sids, gammas = [], []
        for b in self.robin:
            sid = self.sid[id(b)]
            if b._include_force:
                xy = self.points[id(b)]
                U = np.asarray(b.get_Uinc(xy[0].ravel(), xy[1].ravel(), k0), dtype=np.complex128)
                self.ctx.surface_set_U(sid, U.reshape(3, 6, self.ntri[id(b)]))
            if b._include_stiff:
                sids.append(sid)
                gammas.append(complex(b.get_gamma(k0)))
        self.ctx.form_A(k0, sids, gammas)
        return k0

    def solve_point(self, freq, keep_fields=False, raise_on_fail=True, out_bufs=None):
        """One frequency point: A(f), one solve per port, S-parameters.  Returns (S (P,P), stats list, fields dict).
        out_bufs: optional {port_number: preallocated (pinned) complex128[n_field]} receiving the fields."""
        ports = self.ports
        S = np.zeros((len(ports), len(ports)), dtype=np.complex128)
        stats, fields = [], {}
        k0 = self.assemble_frequency(freq)
        want = keep_fields or (out_bufs is not None)
        lock = self.lockstep if self.solver_opts.get("method", "cocr") == "cocr" else 1
        modes = {}
        ja0 = 0
        while ja0 < len(ports):
            group = ports[ja0:ja0 + max(1, min(4, lock))]       # lockstep groups of up to 4 ports
            outs = [out_bufs.get(p.port_number) if out_bufs else None for p in group]
            xs, infos = self.ctx.solve_multi([self.sid[id(p)] for p in group], want_x=want, raise_on_fail=raise_on_fail,
                                             outs=outs, **self.solver_opts)
            ri = self.ctx.recycle_info() if self.recycle else None
            for k, pa in enumerate(group):
                ja = ja0 + k
                info = infos[k]
                info.update(freq=float(freq), port=pa.port_number, lockstep=len(group))
                if ri is not None:
                    info.update(recycled=ri["n"], proj_relres=ri["last_proj_relres"])
                stats.append(info)
                if keep_fields:
                    fields[pa.port_number] = xs[k]
                self.ctx.select_solution(k)
                E_all = self._interp_all(None)       # one device call per solution for the sample points of all ports
                pa.active = True
                _, pout = self._s_data(pa, k0, None, E_all, modes)
                for ib, pb in enumerate(ports):
                    pf, _ = self._s_data(pb, k0, None, E_all, modes)
                    S[ib, ja] = pf / pout
                pa.active = False
            ja0 += len(group)
        return S, stats, fields

    def run(self, freqs, keep_fields=False, raise_on_fail=True, order=None, out_bufs=None, on_point=None) -> SweepResult:
        """Solves every frequency point; results are returned in the order of `freqs` (as emfreq3d.py:658 does).
        order: processing order (list of indices); default hierarchical when recycling is on, else as given.
        out_bufs: optional {port_number: complex128[n_field] host buffer (pinned)} receiving every solved field (D2H each
        point, like data._fields[port] = x at emfreq3d.py:699); on_point(i, S_i, stats) is called after every point."""
        if not self._setup_done:
            self.f_ref = float(np.median(np.asarray(freqs, dtype=float)))
            self.setup()
        ports = self.ports
        pn = [p.port_number for p in ports]
        S = np.zeros((len(freqs), len(ports), len(ports)), dtype=np.complex128)
        res = SweepResult(np.asarray(freqs, dtype=float), pn, S)
        for p in ports:
            p.active = False
        if order is None:
            order = hierarchical_order(len(freqs)) if self.recycle else list(range(len(freqs)))
        stats = {}
        overlap = out_bufs is not None and on_point is None and not keep_fields
        if overlap:                 # nobody reads the buffers before run() returns: their D2H copies overlap the next point
            self.ctx.fields_async(True)
        try:
            for i in order:
                S[i], st, fl = self.solve_point(freqs[i], keep_fields, raise_on_fail, out_bufs=out_bufs)
                stats[i] = st
                if on_point is not None:
                    on_point(i, S[i], st)
                for k, v in fl.items():
                    res.fields[(i, k)] = v
        finally:
            if overlap:
                self.ctx.fields_async(False)        # waits for the last copy
        for i in range(len(freqs)):
            res.stats.extend(stats.get(i, []))
        res.timings = dict(self.timings)
        return res
