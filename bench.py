#!/usr/bin/env python
"""bench.py - frequency points/s of the EMerge frequency-domain hot path on B200 (BASELINE.json metric).

Workload (BASELINE config 4): synthetic ~1M-tet WR-90 rectangular waveguide, 2 RectangularWaveguide ports + PEC walls,
201-point sweep 8-12 GHz.  A "step" is one frequency point of that sweep in the order the sweep driver processes it:
A(f) formation, one solve per port (subspace recycling + preconditioned COCR, true residual <= rtol in FP64),
S-parameter extraction.  The K/M assembly (element kernel + deterministic reduction) runs ONCE at the start of the timed
region, as in the reference (assembler.py:324-331 caches E, B).
  * W warm-up points are solved first and then the recycled subspace is RESET, so the K timed points start cold:
    with no flags K is the whole 201-point sweep (all points of the rank's block), i.e. `value` is the sweep's true
    average throughput; a small --steps measures the expensive first points only (conservative).
  * `value`: inputs resident in HBM (mesh/materials uploaded, patterns built).  `e2e`: a fresh sweep object through the host
    API with HOST buffers - mesh + material upload (pinned), symbolic phase, auxiliary-space setup, assembly, every
    point, and the D2H copy of both solved fields per point into pinned memory - all inside the timed region.  The e2e
    leg runs first (after a kernel warm-up on a small mesh), the resident leg second, each on its own sweep object.
  * N GPUs: one process per GPU, K/M replicated; the points that build the reduced basis (global bisection order) are
    dealt out round-robin in seed rounds, NCCL moves the directions they add between the ranks, then every rank fills
    its contiguous frequency block; the S-parameters are summed over the ranks (emerge_b200/distributed.py).
  * With --steps K smaller than the job, `value` / `e2e` time the first K points of every rank (the expensive cold ones)
    and a second pass times the WHOLE 201-point job: the `full_sweep` block (resident and end-to-end) is the number to
    read strong scaling from.  `roofline` is the dominant kernel (operator application of the block COCR iteration);
    `roofline_assembly` (K+M numeric phase, 9,722 algorithmic B/tet) and `roofline_spmv_c128` (single right-hand side
    complex128 A(f) x, 20 nnz + 36 N B) are BASELINE's other two metrics.  `same_size_as_cpu_sample` runs this GPU path
    on exactly the mesh and points the CPU baseline / reference arm uses: the only like-for-like ratio.

`--impl reference` times the reference's CPU path for the same metric on the host cores: the reference is pure
Python + numba + SciPy and does not exist on the GPU box, so its CPU port (oracle/) runs it: closed-form element
matrices, SciPy COO->CSR, K - k0^2 M + B_p, RCM + SuperLU direct solve (the reference's own fallback solver,
fem/solver.py:535-573), on a bounded sub-sampled mesh of the same waveguide.
"""
from __future__ import annotations

import argparse
import json
import os

# Load every kernel of the process when the CUDA context is created instead of at its first launch: lazy loading cost the
# first sweep of a process ~6 s at 1M tets (first launches inside the solves and CUDA-graph captures: tools/e2e_probe.py,
# 18.9 s -> 12.3 s for the first 12 points).  Must be set before the CUDA driver initialises; a user's own setting wins.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = "freq_points_per_s"
UNIT = "freq-points/s"
A_WG, B_WG = 22.86e-3, 10.16e-3
FREQS = np.linspace(8e9, 12e9, 201)


SLABS = False          # --workload slabs: BASELINE config 5 (ceramic slabs eps_r = 9.8 (1 - 1e-4 j), periodic in z)


def make_waveguide(nx, ny, nz, device=None, slabs=None):
    """device: build the topology tables on that GPU (csrc/topology.cu) instead of numpy on the host (same tables)"""
    from emerge_b200.synthmesh import box_mesh, mesh_tables, tri_ids_of
    from emerge_b200 import bc as B
    L = nz * A_WG / nx
    slabs = SLABS if slabs is None else slabs
    vol = None
    if slabs:
        hz = L / nz                      # layers 5 and 6 of every twelve cell layers are ceramic; the port planes stay in vacuum

        def vol(x, y, z):
            layer = np.floor(z / hz).astype(np.int64) % 12
            return np.where((layer == 5) | (layer == 6), 2, 1)
    box = box_mesh(nx, ny, nz, A_WG, B_WG, L, vol_fn=vol, node_order=os.environ.get("EMB_MESH_ORDER", "lex"))
    if device is None:
        t = mesh_tables(box.nodes_xyz, box.tets)
    else:
        from emerge_b200.lib import Context
        tc = Context(device)
        t = tc.mesh_tables(box.nodes_xyz, box.tets)
        tc.close()
    nT = t.tets.shape[1]
    er = np.zeros((3, 3, nT), complex)
    er[0, 0] = er[1, 1] = er[2, 2] = 1
    ur = er.copy()
    if slabs:
        cer = box.tet_vol == 2
        for k in range(3):
            er[k, k, cer] = 9.8 * (1 - 1e-4j)
    tag = lambda k: tri_ids_of(t, box.face_tris[box.face_tag == k])
    bcs = [B.PEC(np.concatenate([tag(k) for k in (1, 2, 3, 4)])),
           B.RectangularWaveguide(tag(5), 1, B.CoordSys(origin=(0, 0, 0)), (A_WG, B_WG)),
           B.RectangularWaveguide(tag(6), 2, B.CoordSys(origin=(0, 0, L)), (A_WG, B_WG))]
    return box, t, er, ur, bcs, L


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.samples, self.stop = index, [], False
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([v.strip() for v in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=6)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if len(s) >= 6 and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) >= 6 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples if len(s) >= 6 for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------ CPU port
def cpu_port_setup(nx, ny, nz):
    """Reference CPU path on a bounded sample (oracle port): returns a closure running one frequency point."""
    import scipy.sparse.linalg as spla
    from scipy.sparse.csgraph import reverse_cuthill_mckee
    from oracle import ned2_oracle as O
    box, t, er, ur, bcs, L = make_waveguide(nx, ny, nz)
    N = t.n_field
    state = {}

    def assemble():
        E, Bm = O.assemble_EB(t.nodes, t.tets, t.edges, t.tris, t.edge_lengths, t.tet_to_field, t.tet_to_edge, ur, er)
        state["E"], state["B"] = E, Bm
    pec = np.unique(np.asarray(t.tri_to_field)[:, bcs[0].tri_ids].ravel())
    solve_ids = np.setdiff1d(np.arange(N), pec)
    DP = O.dunavant4()
    surf = []
    for p in bcs[1:]:
        ids = p.tri_ids
        v = t.nodes[:, t.tris[:, ids]]
        loc = np.einsum("ij,jvn->ivn", p.get_inv_basis(), v - p.cs.origin[:, None, None])
        x, y = loc[0].T, loc[1].T
        surf.append((p, ids, x, y, O.gen_csr_tri(N, t.tri_to_field, ids, O.tri_surface_matrix(x, y))))

    def point(f):
        k0 = 2 * np.pi * f / 299792458
        K = (state["E"] - state["B"] * k0 ** 2).tocsr()                      # assembler.py:333
        rhs = []
        for p, ids, x, y, S in surf:
            K = K + complex(p.get_gamma(k0)) * S                             # assembler.py:383
            xq, yq = x @ DP[1:4], y @ DP[1:4]
            U = p.get_Uinc(xq.T.ravel(), yq.T.ravel(), k0).reshape(3, 6, len(ids))
            bv = np.zeros(N, dtype=complex)
            np.add.at(bv, t.tri_to_field[:, ids].T, O.tri_forcing(x, y, U[:2]))
            rhs.append(bv)
        Asel = K.tocsc()[np.ix_(solve_ids, solve_ids)]                       # solver.py:434
        perm = reverse_cuthill_mckee(Asel.tocsr(), symmetric_mode=False)     # solver.py:146
        Asort = Asel[perm][:, perm].tocsc()
        lu = spla.splu(Asort, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.01,
                       options=dict(SymmetricMode=True))                     # solver.py:269
        out = []
        for bv in rhs:
            xs = lu.solve(bv[solve_ids][perm])
            out.append(xs)
        return out
    return t, assemble, point


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nx, ny, nz = args.ref_cells
    if args.steps <= 0:
        args.steps = 2
    t0 = time.perf_counter()
    t, assemble, point = cpu_port_setup(nx, ny, nz)
    assemble()
    setup_s = time.perf_counter() - t0
    fr = FREQS
    for i in range(args.warmup):
        point(fr[i % len(fr)])
    t0 = time.perf_counter()
    for i in range(args.steps):
        point(fr[(args.warmup + i) % len(fr)])
    dt = time.perf_counter() - t0
    val = args.steps / dt
    cores = os.cpu_count()
    sample = (f"{t.tets.shape[1]}-tet / {t.n_field}-dof sub-sampled WR-90 waveguide ({nx}x{ny}x{nz} cells), per step: "
              f"K(f) formation + RCM + SuperLU factorisation + 2 port solves (SciPy, 1 thread by construction: "
              f"SuperLU is serial); assembly ({setup_s:.1f}s) outside the steps as the reference caches E,B")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": reference_config(args, t),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "host_cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def reference_config(args, t):
    """what the reference arm actually ran: NOT the GPU arm's mesh (the direct solve is infeasible there)"""
    nx, ny, nz = args.ref_cells
    gx, gy, gz = args.cells
    return {"workload": f"SCALED-DOWN sample of BASELINE config 4: synthetic WR-90 rectangular waveguide, {nx}x{ny}x{nz} cells x 6 "
                        f"Kuhn tets = {6*nx*ny*nz} tets (the GPU arm runs {6*gx*gy*gz} tets), 2 RectangularWaveguide ports + PEC "
                        f"walls, points of the 201-point sweep 8-12 GHz; step = one frequency point (K(f) + RCM + SuperLU "
                        f"factorisation + 2 port solves), K/M assembly once per job",
            "cells": [nx, ny, nz], "ref_cells": [nx, ny, nz], "tets": int(t.tets.shape[1]), "n_field": int(t.n_field),
            "gpu_arm_cells": [gx, gy, gz], "workload_scaled_down": True, "same_config_as_gpu_arm": False,
            "solver": "RCM + SuperLU (scipy.sparse.linalg.splu), the reference's ParallelRoutine (fem/solver.py:535-566)",
            "note": "ratios against the GPU arm's 1M-tet line are NOT like-for-like; use the GPU line's "
                    "same_size_as_cpu_sample block (same mesh, same points)"}


def workload_config(args):
    nx, ny, nz = args.cells
    what = ("dielectric-loaded WR-90 waveguide (ceramic slabs eps_r = 9.8 (1 - 1e-4 j), two of every twelve cell layers; BASELINE "
            "config 5)" if SLABS else "WR-90 rectangular waveguide (BASELINE config 4)")
    return {"workload": f"synthetic {what}, {nx}x{ny}x{nz} cells x 6 Kuhn tets = {6*nx*ny*nz} tets, "
                        f"2 RectangularWaveguide ports + PEC walls, {len(FREQS)}-point sweep 8-12 GHz; "
                        f"step = one frequency point (A(f) + 2 port solves + S-parameters), K/M assembly once per job",
            "cells": [nx, ny, nz], "rtol": args.rtol,
            "solver": "reduced-basis recycling across points (affine A(f)) + block COCR over the ports (one Krylov space) on the "
                      "complex64 symmetric part (FP64 vectors and arithmetic) / FP64 defect correction on A(f), "
                      "additive multilevel (Hiptmair-Xu + smoothed-aggregation AMG) preconditioner, iteration replayed from a CUDA graph",
            "precision": "complex128 arithmetic, vectors, A(f) and residuals; only the VALUES of the inner (preconditioned) operator As are stored complex64",
            "recycle_vectors": args.recycle, "snapshot_rtol_factor": args.snap, "coarse_basis_preconditioner": not args.no_coarse_basis,
            "order": "seed rounds over the global bisection order (round-robin over the ranks), then bisection order within each rank's frequency block",
            "l2_policy": "inputs larger than L2 (A(f) alone is 5.3 GB at 1M tets)", "parallelism": f"freq-block x{args.gpus}"}


def e2e_pass(args, torch, dist, rank, world, local, t, er, ur, bcs, max_points, barrier):
    """End-to-end leg: a fresh sweep object through the host API, host buffers in pinned memory; everything from the
    construction of the sweep object to the last solved field on the host is inside the timed region."""
    from emerge_b200.sweep import FrequencySweep
    from emerge_b200.distributed import ShardedSweep
    N = t.n_field
    er_p = torch.from_numpy(er).pin_memory().numpy()
    ur_p = torch.from_numpy(ur).pin_memory().numpy()
    outs = {p.port_number: torch.empty(N, dtype=torch.complex128).pin_memory().numpy() for p in bcs[1:]}
    barrier()
    t0 = time.perf_counter()
    sw2 = FrequencySweep(t, er_p, ur_p, bcs, device=local, recycle=args.recycle, recycle_snap=args.snap,
                         coarse_basis=not args.no_coarse_basis)
    sw2.solver_opts.update(rtol=args.rtol, precond=args.precond)
    sw2.f_ref = float(np.median(FREQS))
    sw2.setup()
    sh2 = ShardedSweep(sw2, FREQS, rank, world, dist=dist, device=local)
    for p in sw2.ports:
        p.active = False
    res = sh2.run(max_points=max_points, out_bufs=outs, raise_on_fail=False)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3     # host wall clock: the region contains host work (setup) by design
    barrier()
    npts = max(1, len(res.solved))
    mesh_bytes = (np.asarray(t.nodes).nbytes + np.asarray(t.tets).nbytes + np.asarray(t.tris).nbytes
                  + np.asarray(t.tet_to_field).nbytes + np.asarray(t.tri_to_field).nbytes)
    per_step_h2d = sum(18 * 16 * sw2.ntri[id(p)] for p in sw2.ports)
    h2d = (er_p.nbytes + ur_p.nbytes + mesh_bytes) / npts + per_step_h2d
    # fields of every port + per solved port one read-back of E (3 complex128) at the S-parameter sample points of all ports
    d2h = sum(o.nbytes for o in outs.values()) + sum(3 * 16 * sw2._sp[id(p)]["pts"].shape[1] * len(sw2.ports) for p in sw2.ports)
    out = dict(ms=ms, points=len(res.solved), h2d=h2d, d2h=d2h, setup=dict(sw2.timings), split=dict(sh2.timings),
               not_converged=res.not_converged, max_relres=res.max_relres)
    sw2.ctx.close()
    return out


def resident_pass(args, sw, dist, rank, world, local, max_points, barrier, clock=True):
    """Resident leg: mesh, materials, patterns and auxiliary spaces already on the device; K/M assembly + the points,
    from an EMPTY reduced basis, timed with CUDA events on the library's stream."""
    from emerge_b200.distributed import ShardedSweep
    ctx = sw.ctx
    sh = ShardedSweep(sw, FREQS, rank, world, dist=dist, device=local)
    for p in sw.ports:
        p.active = False
    ctx.recycle_config(args.recycle, args.snap)      # forget everything earlier passes left in the basis
    ctx.spmv_sampled()
    ctx.precond_sampled()
    barrier()
    l0, g0 = ctx.launches, ctx.graph_launches
    cs = ClockSampler(local) if clock else None
    if cs:
        cs.__enter__()
    ctx.timer_start()
    ctx.assemble_KM()
    res = sh.run(max_points=max_points, raise_on_fail=False)
    ms = ctx.timer_stop()
    if cs:
        cs.__exit__()
    barrier()
    spmv_ms, spmv_cnt = ctx.spmv_sampled()
    prec_ms, prec_cnt = ctx.precond_sampled()
    return dict(ms=ms, res=res, S=sh.gather_S(res), launches=ctx.launches - l0, graph_iters=ctx.graph_launches - g0,
                spmv_ms=spmv_ms, spmv_cnt=spmv_cnt, prec_ms=prec_ms, prec_cnt=prec_cnt, split=dict(sh.timings),
                asm={"tet_kernel_ms": ctx.last_ms("tet_kernel"), "reduce_ms": ctx.last_ms("reduce")},
                clocks=cs.summary() if cs else None, rinfo=ctx.recycle_info())


def _iter_stats(stats):
    by_freq = {}
    for s_ in stats:                  # the ports of a lockstep group share one iteration count
        by_freq[s_["freq"]] = max(by_freq.get(s_["freq"], 0), s_["iters"])
    return by_freq


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        import datetime
        # a short collective timeout: a mismatched collective must not hold the GPU box for the default 10 minutes
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=240))
    from emerge_b200.sweep import FrequencySweep
    nx, ny, nz = args.cells
    t0 = time.perf_counter()
    box, t, er, ur, bcs, L = make_waveguide(nx, ny, nz, device=None if args.host_tables else local)
    host_mesh_s = time.perf_counter() - t0
    dev = f"cuda:{local}"

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allsum(v):
        x = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(x, op=dist.ReduceOp.SUM)
        return float(x.item())

    def allmax(v):
        x = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(x, op=dist.ReduceOp.MAX)
        return float(x.item())

    def allgather_obj(o):
        if dist is None:
            return [o]
        out = [None] * world
        dist.all_gather_object(out, o)
        return out

    nf = len(FREQS)
    partial = 0 < args.steps < -(-nf // world)             # --steps K smaller than a rank's share of the job
    K = args.steps if partial else None
    do_full = (not partial) or (not args.no_full_sweep)
    do_e2e = args.e2e_steps >= 0
    # kernel warm-up on a small mesh (CUDA context, module load, every kernel of the path once), then the end-to-end legs
    # FIRST: they have to see the process as a user's script would (a device heap that just released tens of GB makes
    # cudaMalloc ten times slower, which is what the e2e leg measured when it ran after the resident leg)
    wbox, wt, wer, wur, wbcs, _ = make_waveguide(8, 4, 12)
    wsw = FrequencySweep(wt, wer, wur, wbcs, device=local, recycle=args.recycle, recycle_snap=args.snap,
                         coarse_basis=not args.no_coarse_basis)
    wsw.solver_opts.update(rtol=args.rtol, precond=args.precond)
    wsw.run(list(FREQS[:: max(1, nf // max(1, args.warmup))][:max(3, args.warmup)]), raise_on_fail=False)
    wsw.ctx.close()
    del wsw
    e2e_k = e2e_pass(args, torch, dist, rank, world, local, t, er, ur, bcs, K, barrier) if (do_e2e and partial) else None
    e2e_full = e2e_pass(args, torch, dist, rank, world, local, t, er, ur, bcs, None, barrier) if (do_e2e and do_full) else None

    sw = FrequencySweep(t, er, ur, bcs, device=local, recycle=args.recycle, recycle_snap=args.snap,
                        coarse_basis=not args.no_coarse_basis)
    sw.solver_opts.update(rtol=args.rtol, precond=args.precond)
    sw.f_ref = float(np.median(FREQS))
    t0 = time.perf_counter()
    sw.setup()
    setup_s = time.perf_counter() - t0
    ctx = sw.ctx
    nnz_s, Ns, N = int(ctx.lib.emb_csr_nnz(ctx.h, 2)), ctx.n_solve, ctx.n_field
    nnz_full = int(ctx.lib.emb_nnz(ctx.h))
    # warm-up: W points on the full-size operator (graph capture, allocations), forgotten by the reset in resident_pass
    for p in sw.ports:
        p.active = False
    from emerge_b200.sweep import hierarchical_order
    for i in hierarchical_order(nf)[:args.warmup]:
        sw.solve_point(FREQS[i], raise_on_fail=False)
    main = resident_pass(args, sw, dist, rank, world, local, K, barrier)
    full = resident_pass(args, sw, dist, rank, world, local, None, barrier, clock=False) if (partial and do_full) else main
    free_b, total_b = torch.cuda.mem_get_info(local)
    hbm = allgather_obj({"rank": rank, "used_GB": (total_b - free_b) / 1e9, "total_GB": total_b / 1e9})
    # BASELINE's third metric: the single right-hand side complex128 operator application A(f) x, timed alone
    sw.assemble_frequency(float(FREQS[nf // 2]))
    spmv128_ms = ctx.spmv_bench(20, nv=1, fp32=False)

    def job(p, e):
        """whole-job numbers of a pass: points of all ranks / slowest rank's time"""
        pts = allsum(len(p["res"].solved))
        ms = allmax(p["ms"])
        out = {"points": int(pts), "value": pts / (ms / 1e3), "ms": ms}
        if e is not None:
            epts, ems = allsum(e["points"]), allmax(e["ms"])
            out.update(e2e_value=epts / (ems / 1e3), e2e_ms=ems, e2e_points=int(epts))
        return out
    jm = job(main, e2e_k if partial else e2e_full)
    jf = job(full, e2e_full) if full is not main else jm
    splits = allgather_obj({"rank": rank, **{k: round(v, 4) if isinstance(v, float) else v for k, v in full["split"].items()},
                            "points": len(full["res"].solved), "device_ms": full["ms"],
                            "iterating_points": sum(1 for v in _iter_stats(full["res"].stats).values() if v > 0),
                            "iterations": int(sum(_iter_stats(full["res"].stats).values())),
                            "host_setup_s": e2e_full["setup"].get("aux_setup_s") if e2e_full else None,
                            "e2e_ms": e2e_full["ms"] if e2e_full else None})
    stats_all = [s_ for part in allgather_obj(main["res"].stats) for s_ in part]
    launches_total = allsum(main["launches"])
    if rank == 0:
        peak, peak_src = peaks()
        nv = min(4, len(sw.ports)) if sw.lockstep > 1 else 1
        nv = 4 if nv == 3 else nv
        # dominant kernel: the operator application of the block COCR iteration on complex64 2x2 blocks - k_bsell_tma<2> (SELL
        # layout, bulk-async staged) for two right-hand sides, k_bspmv (block-CSR) otherwise: per nonzero 8 B value + 1 B (one
        # 4 B column per 2x2 block), per row 4 B rowptr (8 B per block-row) + NV x (16 B x + 16 B y)   
        spmv_bytes = 9 * nnz_s + (4 + 32 * nv) * Ns + 8
        spmv_ms = main["spmv_ms"]
        achieved = spmv_bytes / (spmv_ms * 1e-3) / 1e9 if spmv_ms > 0 else None
        traffic = asm_traffic = None
        tp = os.path.join(REPO, "profiles", "spmv_traffic.json")
        if os.path.exists(tp) and os.path.getsize(tp) > 0 and (nx, ny, nz) == (44, 20, 190):
            tj = json.load(open(tp))
            if tj.get("nv") == nv and tj.get("values") == "complex64":
                traffic = tj.get("dram_bytes_per_launch")
            asm_traffic = tj.get("assembly_dram_bytes")
        S_all = main["S"]
        solved = sorted({int(np.argmin(np.abs(FREQS - s_["freq"]))) for s_ in stats_all})
        S21 = np.abs(S_all[solved, 1, 0])
        by_freq = _iter_stats(stats_all)
        iters = list(by_freq.values())
        asm = main["asm"]
        asm_ms = asm["tet_kernel_ms"] + asm["reduce_ms"]
        nT = int(t.tets.shape[1])
        asm_bytes = 9722 * nT
        spmv128_bytes = 20 * nnz_s + 36 * Ns + 4
        line = {"metric": METRIC, "value": jm["value"], "unit": UNIT, "n_gpus": world,
                "steps": args.steps if partial else jm["points"], "warmup": args.warmup,
                "ms_per_step": jm["ms"] / (args.steps if partial else max(1, -(-jm["points"] // world))), "higher_is_better": True,
                "scaling": "weak" if partial else "strong", "vs_baseline": None,
                # with --steps K below a rank's share, `value` times the first K points of EVERY rank: cold seed points at
                # N = 1, mostly points filled from the shared basis at large N - different work, so value(N) / value(1) is
                # not a scaling efficiency; the whole-job numbers of `full_sweep` are
                "value_comparable_across_n": not partial,
                "strong_scaling_value": jf["value"], "strong_scaling_e2e": jf.get("e2e_value"),
                "dtype": "f64", "data": "synthetic", "config": workload_config(args),
                "e2e": {"value": jm.get("e2e_value"), "unit": UNIT,
                        "h2d_bytes_per_step": int((e2e_k or e2e_full)["h2d"]) if (e2e_k or e2e_full) else None,
                        "d2h_bytes_per_step": int((e2e_k or e2e_full)["d2h"]) if (e2e_k or e2e_full) else None,
                        "points": jm.get("e2e_points"),
                        "timed": "host wall clock around FrequencySweep() construction, setup() and the points"},
                "full_sweep": {"points": jf["points"], "value": jf["value"], "ms": jf["ms"], "e2e_value": jf.get("e2e_value"),
                               "e2e_ms": jf.get("e2e_ms"), "unit": UNIT, "scaling": "strong",
                               "what": "the whole 201-point job (assembly + every point from an empty basis) on all ranks: "
                                       "points / slowest rank's time; this is the strong-scaling number",
                               "per_rank": splits},
                "gpu_launches": int(launches_total),
                "roofline": {"kernel": (f"k_bsell_tma<8,24,2,1> (operator application of the block COCR iteration: complex64 2x2 "
                                        f"blocks in a SELL-8 layout, slice streams staged with cp.async.bulk + mbarrier under an L2 "
                                        f"evict-first policy, {nv} interleaved right-hand sides)"
                                        if nv == 2 else
                                        f"k_bspmv<NV={nv}, complex64 values> (2x2 block-CSR operator application of the block "
                                        f"COCR iteration on {nv} interleaved right-hand sides)"),
                             "bound": "hbm", "achieved": achieved,
                             "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                             "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                             "algorithmic_bytes_per_launch": spmv_bytes, "avg_launch_ms": spmv_ms, "sampled_launches": main["spmv_cnt"]},
                "roofline_assembly": {"kernel": "k_tet_records + k_asm_rows (fused numeric phase: K and M values written once)",
                                      "bound": "hbm", "achieved": asm_bytes / (asm_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                      "frac": asm_bytes / (asm_ms * 1e-3) / 1e9 / peak, "traffic": asm_traffic,
                                      "algorithmic_bytes_per_launch": asm_bytes, "avg_launch_ms": asm_ms,
                                      "Mtet_per_s": nT / (asm_ms * 1e3), **asm},
                "roofline_spmv_c128": {"kernel": "k_bspmv<NV=1, complex128 values> (A(f) x, one right-hand side: SURVEY 8d)",
                                       "bound": "hbm", "achieved": spmv128_bytes / (spmv128_ms * 1e-3) / 1e9, "peak": peak,
                                       "unit": "GB/s", "frac": spmv128_bytes / (spmv128_ms * 1e-3) / 1e9 / peak,
                                       "algorithmic_bytes_per_launch": spmv128_bytes, "avg_launch_ms": spmv128_ms},
                "assembly": {**asm, "Mtet_per_s": nT / (asm_ms * 1e3), "form_A_ms": ctx.last_ms("form_A")},
                "solver": {"lockstep_iterations_total": int(np.sum(iters)), "lockstep_width": nv,
                           "points_that_iterated": int(sum(1 for v in iters if v > 0)),
                           "iterations_per_iterating_point": float(np.sum(iters) / max(1, sum(1 for v in iters if v > 0))),
                           "points": jm["points"], "max_iters_per_point": int(max(iters)), "rtol": args.rtol,
                           "precond_apply_ms": main["prec_ms"], "precond_samples": main["prec_cnt"],
                           "graph_replayed_iterations": int(main["graph_iters"]),
                           "max_relres": float(max(s_["relres"] for s_ in stats_all)),
                           "not_converged": int(sum(1 for s_ in stats_all if not s_.get("converged", True))),
                           "recycled_directions": main["rinfo"]["n"], "recycle_term_products": main["rinfo"]["spmvs"],
                           "abs_S21_minmax": [float(S21.min()), float(S21.max())]},
                "sizes": {"tets": nT, "n_field": N, "n_solve": Ns, "nnz_full": nnz_full, "nnz_solve": nnz_s},
                "setup": {"host_mesh_tables_s": host_mesh_s, "gpu_setup_s": setup_s, **{k: v for k, v in sw.timings.items()}},
                "e2e_setup": (e2e_k or e2e_full)["setup"] if (e2e_k or e2e_full) else {},
                "e2e_split": {"steps_leg": e2e_k["split"] if e2e_k else None, "full_leg": e2e_full["split"] if e2e_full else None,
                              "steps_leg_ms": e2e_k["ms"] if e2e_k else None, "full_leg_ms": e2e_full["ms"] if e2e_full else None},
                "hbm_after_sweep": hbm, "clocks": main["clocks"]}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args)
            line["same_size_as_cpu_sample"] = gpu_same_size(args, local, line["cpu_baseline"])
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


CPU_SAMPLE_FREQS = (8.0e9, 9.0e9)      # the points the CPU baseline solves (indices 0 and 50 of the 201-point sweep)


def gpu_same_size(args, device, cpu):
    """This GPU path on exactly the mesh AND the points the CPU baseline uses (the only like-for-like comparison), plus
    the whole 201-point sweep on that mesh."""
    from emerge_b200.sweep import FrequencySweep
    nx, ny, nz = args.ref_cells
    box, t, er, ur, bcs, L = make_waveguide(nx, ny, nz)
    sw = FrequencySweep(t, er, ur, bcs, device=device, recycle=args.recycle, coarse_basis=not args.no_coarse_basis)
    sw.solver_opts.update(rtol=args.rtol, precond=args.precond)
    sw.f_ref = float(np.median(FREQS))
    sw.setup()
    for p in sw.ports:
        p.active = False
    for f in FREQS[len(FREQS) // 2:len(FREQS) // 2 + 2]:
        sw.solve_point(f)
    sw.ctx.recycle_config(args.recycle, args.snap)
    pts = [float(f) for f in CPU_SAMPLE_FREQS]
    sw.ctx.timer_start()                       # K/M already assembled, as in the CPU sample
    sw.run(pts, order=list(range(len(pts))))
    ms_pts = sw.ctx.timer_stop()
    sw.ctx.recycle_config(args.recycle, args.snap)
    sw.ctx.timer_start()
    sw.ctx.assemble_KM()
    sw.run(FREQS)
    ms = sw.ctx.timer_stop()
    sw.ctx.close()
    same = len(pts) / (ms_pts / 1e3)
    return {"cells": [nx, ny, nz], "tets": int(t.tets.shape[1]),
            "same_points": {"value": same, "unit": UNIT, "points_GHz": [f / 1e9 for f in pts],
                            "sample": "the CPU baseline's own points from a cold reduced basis, K/M assembled beforehand as in the CPU sample",
                            "ratio_vs_cpu_baseline": same / cpu["value"] if cpu and cpu.get("value") else None},
            "value": len(FREQS) / (ms / 1e3), "unit": UNIT,
            "sample": "whole 201-point sweep (assembly included) from a cold recycled subspace"}


def cpu_baseline(args):
    """Oracle port (kind "port") on rank 0's host cores, bounded sample of the same workload."""
    nx, ny, nz = args.ref_cells
    t0 = time.perf_counter()
    t, assemble, point = cpu_port_setup(nx, ny, nz)
    assemble()
    asm_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    for f in CPU_SAMPLE_FREQS:
        point(f)
    n = len(CPU_SAMPLE_FREQS)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": 1, "host_cores": os.cpu_count(), "kind": "port",
            "cells": [nx, ny, nz], "tets": int(t.tets.shape[1]),
            "assembly_Mtet_per_s": t.tets.shape[1] / asm_s / 1e6,
            "sample": f"{t.tets.shape[1]}-tet / {t.n_field}-dof sub-sampled waveguide ({nx}x{ny}x{nz} cells), {n} frequency points: "
                      f"K(f) + RCM + SuperLU + 2 solves per point (SciPy; SuperLU is serial); oracle assembly {asm_s:.1f}s "
                      f"({t.tets.shape[1]/asm_s/1e6:.4f} Mtet/s, numpy port) not included"}


REF_CANDIDATES = [(6, 3, 10), (8, 4, 16), (10, 5, 20), (12, 6, 24), (12, 6, 30), (14, 7, 30), (16, 8, 36)]


def pick_ref_cells(total_steps, budget_s=150.0):
    """Largest sub-sampled waveguide whose SuperLU steps fit the time budget: step ~ 27.7 s x (tets/12960)^2.3
    (27.7 s per frequency at 12,960 tets on this class of host)."""
    best = REF_CANDIDATES[0]
    for c in REF_CANDIDATES:
        nT = 6 * c[0] * c[1] * c[2]
        if total_steps * 27.7 * (nT / 12960.0) ** 2.3 <= budget_s:
            best = c
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="timed frequency points per rank; 0 = the rank's whole block")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="emerge_b200", choices=["emerge_b200", "reference"])
    ap.add_argument("--cells", type=lambda s: tuple(int(v) for v in s.split(",")), default=(44, 20, 190))
    ap.add_argument("--ref-cells", type=lambda s: tuple(int(v) for v in s.split(",")), default=None)
    ap.add_argument("--rtol", type=float, default=1e-8)
    ap.add_argument("--precond", default="multilevel")
    ap.add_argument("--e2e-steps", type=int, default=0, help="points of the end-to-end pass; 0 = same as --steps, -1 = skip")
    ap.add_argument("--recycle", type=int, default=40)
    ap.add_argument("--snap", type=float, default=0.3, help="points that iterate are solved to snap * rtol")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-coarse-basis", action="store_true",
                    help="switch off the reduced basis as an extra coarse space of the preconditioner (default on)")
    ap.add_argument("--no-full-sweep", action="store_true", help="skip the whole-job pass when --steps is smaller than the job")
    ap.add_argument("--workload", default="waveguide", choices=["waveguide", "slabs"],
                    help="waveguide: BASELINE config 4 (default); slabs: config 5 (use with --cells 76,34,323 --points 401)")
    ap.add_argument("--points", type=int, default=201, help="frequency points of the sweep, 8-12 GHz")
    ap.add_argument("--host-tables", action="store_true", help="build the mesh topology tables with numpy on the host")
    ap.add_argument("--mem", action="store_true", help="report the HBM footprint (cudaMemGetInfo) in the line")
    args = ap.parse_args()
    global FREQS, SLABS
    FREQS = np.linspace(8e9, 12e9, args.points)
    SLABS = args.workload == "slabs"
    if args.ref_cells is None:
        args.ref_cells = pick_ref_cells(max(args.steps, 1) + args.warmup) if args.impl == "reference" else pick_ref_cells(3, 60.0)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
