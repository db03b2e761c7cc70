"""Frequency-sharded sweep: one process per GPU, K/M and the eliminated pattern replicated.

Mirror of the reference's only data-parallel path, frequency_domain_par (fem/physics/edm/emfreq3d.py:469-605): there
a multiprocessing.Pool pickles one assembled A(f) per job to each worker (:538-539); here every rank assembles K/M
itself (3 ms at 1M tets; a broadcast would move 9 GB) and only scalars cross the host/device boundary per point.

Schedule (what bounds strong scaling is the chain of COLD solves that build the reduced basis, not a collective):
  1. seed rounds - the points of the GLOBAL bisection order (both band edges, the midpoint, the quarter points, ...) are
     dealt out round-robin, one per rank per round.  These are exactly the points that have to iterate in a single-GPU
     sweep (about 11 of 201 for the 8-12 GHz waveguide band), so W ranks pay ceil(11 / W) cold solves each instead of
     every rank re-discovering the band inside its own block.  After each round every rank contributes the directions
     its point added to its reduced basis (at most one per port); all ranks import all of them (one all_gather of
     [ports, n_solve] complex128 per rank over NCCL / NVLink, a few hundred MB).  The rounds stop when a whole round
     added no direction (every seed point was already represented by the shared basis) or after max_seed_rounds.
  2. fill - every rank solves the not-yet-solved points of its contiguous block (bisection order inside the block) from
     the shared basis; a point that still has to iterate does so locally.
  3. the S-parameter blocks are summed over the ranks (all_reduce, kilobytes) together with the convergence record:
     a point that did not converge raises on EVERY rank after the sweep (never inside it, where it would leave the other
     ranks waiting in a collective).
No collective runs inside a Krylov iteration.
"""
from __future__ import annotations

import numpy as np

from .sweep import SweepResult, hierarchical_order


def block_of(n: int, rank: int, world: int) -> np.ndarray:
    """Contiguous block of rank `rank` when n points are split over `world` ranks (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return np.arange(start, start + base + (1 if rank < rem else 0))


class GpuEngine:
    """What ShardedSweep needs from a FrequencySweep + its device context (the CPU tests substitute a fake).
    host_staged: the exchange buffers live in (pinned) host memory and are staged through one device vector - for
    process groups that cannot move CUDA tensors (gloo's all_gather), e.g. two ranks sharing one GPU in the tests."""

    def __init__(self, sweep, device, host_staged=False):
        import torch
        self.torch = torch
        self.sweep = sweep
        self.device = torch.device("cuda", device)
        self.n_ports = len(sweep.ports)
        self.n = sweep.ctx.n_solve
        self.host_staged = bool(host_staged)
        self._stage = torch.zeros(self.n, dtype=torch.complex128, device=self.device) if self.host_staged else None

    def solve_point(self, f, **kw):
        return self.sweep.solve_point(f, **kw)

    def fields_async(self, on):
        """the D2H copies of the solved fields overlap the next point (emb_fields_async); off = wait for the last one"""
        self.sweep.ctx.fields_async(on)

    def accepted(self):
        """monotonic count of directions this rank's basis has accepted (emb_recycle_accepted)"""
        return self.sweep.ctx.recycle_accepted()

    def new_buffer(self, k):
        if self.host_staged:
            return self.torch.zeros((k, self.n), dtype=self.torch.complex128).pin_memory()
        return self.torch.zeros((k, self.n), dtype=self.torch.complex128, device=self.device)

    def tensor(self, values, dtype=None):
        return self.torch.as_tensor(np.asarray(values), dtype=dtype, device="cpu" if self.host_staged else self.device)

    def export_newest(self, k, buf):
        """copy the k newest recycled directions into rows 0..k-1 of buf; synchronises the library stream"""
        for j in range(k):
            if self.host_staged:
                self.sweep.ctx.recycle_export(j, self._stage.data_ptr())
                buf[j].copy_(self._stage)
            else:
                self.sweep.ctx.recycle_export(j, buf[j].data_ptr())
        if self.host_staged:
            self.torch.cuda.synchronize(self.device)

    def sync(self):
        """the collectives run on torch's stream, the library on its own: one device synchronisation per exchange"""
        self.torch.cuda.synchronize(self.device)

    def import_direction(self, row):
        if self.host_staged:
            self._stage.copy_(row)
            self.torch.cuda.synchronize(self.device)
            self.sweep.ctx.recycle_import(self._stage.data_ptr())
        else:
            self.sweep.ctx.recycle_import(row.data_ptr())


class ShardedSweep:
    def __init__(self, sweep, freqs, rank=0, world=1, dist=None, device=0, engine=None, max_seed_rounds=6,
                 min_seed_rounds=1, host_staged=None):
        self.freqs = np.asarray(freqs, dtype=float)
        self.rank, self.world, self.dist = rank, world, dist
        self.block = block_of(len(self.freqs), rank, world)
        self.parallel = dist is not None and world > 1
        self.max_seed_rounds = max_seed_rounds if self.parallel else 0
        self.min_seed_rounds = min(min_seed_rounds, self.max_seed_rounds)
        if engine is None:
            if host_staged is None:       # gloo cannot all_gather CUDA tensors
                host_staged = bool(self.parallel and dist.get_backend() == "gloo")
            engine = GpuEngine(sweep, device, host_staged=host_staged)
        self.engine = engine
        self.global_order = hierarchical_order(len(self.freqs))
        self.rounds = 0
        self.exchanged = 0
        self.import_s = 0.0
        self.timings = {}
        self._bufs = None

    # ------------------------------------------------------------------ plan
    def seed_point(self, rnd: int):
        """global index of this rank's point in seed round rnd, or None"""
        k = rnd * self.world + self.rank
        return int(self.global_order[k]) if k < len(self.global_order) else None

    def fill_order(self, rounds: int) -> list:
        """this rank's block in bisection order without the points solved (by any rank) in the first `rounds` seed rounds"""
        seeded = set(int(i) for i in self.global_order[:rounds * self.world])
        return [int(self.block[i]) for i in hierarchical_order(len(self.block)) if int(self.block[i]) not in seeded]

    def order(self, rounds: int | None = None) -> list:
        """processing order of this rank for a given number of seed rounds (default: the rounds of the last run)"""
        rounds = self.rounds if rounds is None else rounds
        seeds = [self.seed_point(r) for r in range(rounds)]
        return [i for i in seeds if i is not None] + self.fill_order(rounds)

    # ------------------------------------------------------------------ direction exchange
    def _exchange(self, n_new):
        eng, dist = self.engine, self.dist
        P, W = eng.n_ports, self.world
        if self._bufs is None:
            self._bufs = (eng.new_buffer(P), eng.new_buffer(W * P))
        mine, allb = self._bufs
        n_new = max(0, min(int(n_new), P))
        if n_new:
            eng.export_newest(n_new, mine)
        cnt = eng.tensor([n_new], dtype=None)
        counts = [cnt.clone() for _ in range(W)]
        dist.all_gather(counts, cnt)
        counts = [int(c.item()) for c in counts]
        if sum(counts):
            parts = list(allb.view(W, P, -1).unbind(0))
            dist.all_gather(parts, mine)
            eng.sync()
            import time
            t0 = time.perf_counter()
            for r in range(W):
                if r == self.rank:
                    continue
                for j in range(counts[r]):
                    eng.import_direction(parts[r][j])
                    self.exchanged += 1
            self.import_s += time.perf_counter() - t0     # the rest of exchange_s is waiting for the slowest rank + NCCL
        return int(sum(counts))

    # ------------------------------------------------------------------ the sweep
    def run(self, max_points=None, out_bufs=None, raise_on_fail=True, keep_fields=False, fields_out=None) -> SweepResult:
        """Solves this rank's share of the sweep (at most max_points points: benchmark steps).  Returns a SweepResult whose S
        has one row per GLOBAL frequency (rows of points solved by other ranks are zero until gather_S).
        fields_out: optional dict receiving {(global index, port_number): field} of the points this rank solved."""
        import time
        eng = self.engine
        nf, P = len(self.freqs), eng.n_ports
        S = np.zeros((nf, P, P), dtype=np.complex128)
        stats = {}
        budget = len(self.freqs) if max_points is None else int(max_points)

        def solve(i):
            Si, st, fl = eng.solve_point(self.freqs[i], raise_on_fail=False, out_bufs=out_bufs, keep_fields=keep_fields)
            S[i] = Si
            stats[i] = st
            if fields_out is not None:
                for k, v in fl.items():
                    fields_out[(i, k)] = v

        overlap = out_bufs is not None and fields_out is None and hasattr(eng, "fields_async")
        if overlap:                          # the caller reads the buffers after run(): copies overlap the next point
            eng.fields_async(True)
        t0 = time.perf_counter()
        t_ex = 0.0
        self.rounds = 0
        while self.rounds < self.max_seed_rounds and self.rounds * self.world < nf:
            i = self.seed_point(self.rounds)
            n_before = eng.accepted()
            if i is not None and len(stats) < budget:
                solve(i)
            t1 = time.perf_counter()
            total_new = self._exchange(eng.accepted() - n_before)
            t_ex += time.perf_counter() - t1
            self.rounds += 1
            if self.rounds >= self.min_seed_rounds and total_new == 0:
                break
        t_seed = time.perf_counter() - t0
        t0 = time.perf_counter()
        for i in self.fill_order(self.rounds):
            if len(stats) >= budget:
                break
            solve(i)
        if overlap:
            eng.fields_async(False)          # waits for the last copy
        self.timings = dict(seed_s=t_seed - t_ex, exchange_s=t_ex, fill_s=time.perf_counter() - t0,
                            seed_rounds=self.rounds, imported_directions=self.exchanged, import_s=self.import_s)
        res = SweepResult(self.freqs, [], S)
        for i in sorted(stats):
            res.stats.extend(stats[i])
        res.solved = sorted(stats)
        res.timings = dict(self.timings)
        self._check_converged(res, raise_on_fail)
        return res

    def _check_converged(self, res, raise_on_fail):
        """worst relative residual and number of unconverged (point, port) solves over ALL ranks; raises on every rank"""
        bad = [s for s in res.stats if not s.get("converged", True)]
        worst = max([float(s.get("relres", 0.0)) for s in res.stats], default=0.0)
        if not np.isfinite(worst):
            worst = np.inf
        nbad = len(bad)
        if self.parallel:
            t = self.engine.tensor([float(nbad), 0.0 if not np.isfinite(worst) else worst, float(not np.isfinite(worst))],
                                   dtype=None)
            tmax = t.clone()
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
            self.dist.all_reduce(tmax, op=self.dist.ReduceOp.MAX)
            nbad = int(round(float(t[0].item())))
            worst = np.inf if float(tmax[2].item()) > 0 else float(tmax[1].item())
        res.not_converged, res.max_relres = nbad, worst
        if nbad and raise_on_fail:
            from .lib import NotConverged
            where = ", ".join(f"{s['freq'] / 1e9:.4f} GHz port {s['port']} (relres {s['relres']:.2e})" for s in bad[:4])
            raise NotConverged(f"sharded sweep: {nbad} solve(s) did not converge (worst relres {worst:.2e})"
                               + (f"; on this rank: {where}" if where else ""))

    def gather_S(self, res: SweepResult) -> np.ndarray:
        """all ranks receive the S-parameters of every solved point (each point is solved by exactly one rank, so the sum
        of the zero-padded per-rank arrays is the result)"""
        if not self.parallel:
            return res.S
        t = self.engine.tensor(np.ascontiguousarray(res.S).view(np.float64), dtype=None)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return np.ascontiguousarray(t.cpu().numpy()).view(np.complex128).reshape(res.S.shape)
