"""Frequency-sharded sweep: one process per GPU, contiguous frequency blocks, K/M and the eliminated pattern replicated.

Mirror of the reference's only data-parallel path, frequency_domain_par (fem/physics/edm/emfreq3d.py:469-605): there
a multiprocessing.Pool pickles one assembled A(f) per job to each worker (:538-539); here every rank assembles K/M
itself (cheaper than a broadcast at >70 Mtet/s) and only scalars cross the host/device boundary per point.

The ranks are independent except for two exchanges, both over torch.distributed (NCCL over NVLink on the GPU box, gloo in
the CPU tests):
  * exchange rounds: after each of its first points every rank contributes the directions that point added to its
    reduced basis (at most one per port); all ranks import all of them.  The rounds go on while any rank still added a
    direction (at least `seed_rounds`, at most `max_rounds`): in bisection order the points that have to iterate come
    first on every rank, so the ranks are in step while they exchange and run free afterwards.  The expensive solves of a
    sweep are the ones that build the basis (about a dozen points for the 8-12 GHz waveguide band) - shared this way they
    are paid once per job instead of once per rank;
  * the S-parameter blocks are gathered to every rank at the end (all_gather, kilobytes).
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
    """What ShardedSweep needs from a FrequencySweep + its device context (the CPU tests substitute a fake)."""

    def __init__(self, sweep, device):
        import torch
        self.torch = torch
        self.sweep = sweep
        self.device = torch.device("cuda", device)
        self.n_ports = len(sweep.ports)
        self.n = sweep.ctx.n_solve

    def solve_point(self, f, **kw):
        return self.sweep.solve_point(f, **kw)

    def recycle_count(self):
        return self.sweep.ctx.recycle_info()["n"]

    def new_buffer(self, k):
        return self.torch.zeros((k, self.n), dtype=self.torch.complex128, device=self.device)

    def export_newest(self, k, buf):
        """copy the k newest recycled directions into rows 0..k-1 of buf (device tensor)"""
        for j in range(k):
            self.sweep.ctx.recycle_export(j, buf[j].data_ptr())

    def import_direction(self, row):
        self.torch.cuda.synchronize()
        self.sweep.ctx.recycle_import(row.data_ptr())


class ShardedSweep:
    def __init__(self, sweep, freqs, rank=0, world=1, dist=None, device=0, seed_rounds=2, engine=None, max_rounds=16):
        self.freqs = np.asarray(freqs, dtype=float)
        self.rank, self.world, self.dist = rank, world, dist
        self.block = block_of(len(self.freqs), rank, world)
        self.seed_rounds = seed_rounds if (dist is not None and world > 1) else 0
        self.max_rounds = max(max_rounds, self.seed_rounds) if self.seed_rounds > 0 else 0
        self.rounds = 0
        self.engine = engine if engine is not None else GpuEngine(sweep, device)
        self.exchanged = 0

    def order(self):
        """global indices of this rank's block in processing order"""
        return [int(self.block[i]) for i in hierarchical_order(len(self.block))]

    # ------------------------------------------------------------------ direction exchange
    def _exchange(self, n_new):
        eng, dist = self.engine, self.dist
        P = eng.n_ports
        mine = eng.new_buffer(P)
        n_new = max(0, min(int(n_new), P))
        if n_new:
            eng.export_newest(n_new, mine)
        counts = [None] * self.world
        dist.all_gather_object(counts, n_new)
        bufs = [eng.new_buffer(P) for _ in range(self.world)]
        dist.all_gather(bufs, mine)
        for r in range(self.world):
            if r == self.rank:
                continue
            for j in range(counts[r]):
                eng.import_direction(bufs[r][j])
                self.exchanged += 1
        return int(sum(counts))

    # ------------------------------------------------------------------ the sweep
    def run(self, order=None, out_bufs=None, raise_on_fail=False) -> SweepResult:
        """Solves the points `order` (global indices, default: the whole block).  Returns a SweepResult whose S has one
        row per GLOBAL frequency (rows of other ranks' points are zero until gather_S)."""
        eng = self.engine
        order = self.order() if order is None else list(order)
        S = None
        stats = {}
        step = 0
        exchanging = self.seed_rounds > 0
        while step < len(order) or exchanging:
            n_before = eng.recycle_count() if exchanging else 0
            if step < len(order):
                i = order[step]
                Si, st, _ = eng.solve_point(self.freqs[i], raise_on_fail=raise_on_fail, out_bufs=out_bufs)
                if S is None:
                    S = np.zeros((len(self.freqs),) + Si.shape, dtype=np.complex128)
                S[i] = Si
                stats[i] = st
            if exchanging:
                # every rank sees the same total, so all ranks leave the exchange in the same round
                total_new = self._exchange(eng.recycle_count() - n_before)
                self.rounds += 1
                if self.rounds >= self.max_rounds or (self.rounds >= self.seed_rounds and total_new == 0):
                    exchanging = False
            step += 1
        res = SweepResult(self.freqs, [], S if S is not None else np.zeros((len(self.freqs), 0, 0), dtype=np.complex128))
        for i in sorted(stats):
            res.stats.extend(stats[i])
        res.solved = sorted(stats)
        return res

    def gather_S(self, res: SweepResult) -> np.ndarray:
        """all ranks receive the S-parameters of every solved point (sum of the zero-padded per-rank arrays)"""
        if self.dist is None or self.world == 1:
            return res.S
        parts = [None] * self.world
        self.dist.all_gather_object(parts, res.S)
        return np.sum(parts, axis=0)
