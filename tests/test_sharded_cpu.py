"""Host logic of the frequency-sharded sweep on CPU: block partition, processing order, the direction exchange of the
seeding rounds and the S-parameter gather, with torch.distributed (gloo, world_size 2) and a fake engine in place of the
GPU context (no compute calls; the device path itself is covered by the -m gpu tests)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from emerge_b200.distributed import ShardedSweep, block_of
from emerge_b200.sweep import hierarchical_order


def test_blocks_partition_the_sweep():
    for n, w in ((201, 1), (201, 2), (201, 8), (7, 8), (16, 4)):
        blocks = [block_of(n, r, w) for r in range(w)]
        allidx = np.concatenate(blocks)
        assert np.array_equal(allidx, np.arange(n))
        assert max(len(b) for b in blocks) - min(len(b) for b in blocks) <= 1


def test_hierarchical_order_is_a_permutation_ends_first():
    for n in (1, 2, 3, 5, 25, 201):
        o = hierarchical_order(n)
        assert sorted(o) == list(range(n))
        if n >= 3:
            assert o[:3] == [0, n - 1, (n - 1) // 2]


class FakeEngine:
    """Every solved point adds one direction per port: the vector (f, port, rank-independent)."""

    def __init__(self, n_ports=2, n=5, iterating_points=10 ** 9):
        self.n_ports, self.n = n_ports, n
        self.vecs = []            # newest first
        self.imported = []
        self.iterating_points = iterating_points      # only the first points leave directions (the others are "accepted")
        self.solved = 0

    def solve_point(self, f, raise_on_fail=False, out_bufs=None):
        self.solved += 1
        for p in range(self.n_ports if self.solved <= self.iterating_points else 0):
            self.vecs.insert(0, np.full(self.n, f * 1e-9 + 1j * p))
        S = np.full((self.n_ports, self.n_ports), f * 1e-9, dtype=complex)
        return S, [dict(freq=float(f), port=p, iters=1, relres=0.0) for p in range(self.n_ports)], {}

    def recycle_count(self):
        return len(self.vecs)

    def new_buffer(self, k):
        return torch.zeros((k, self.n), dtype=torch.complex128)

    def export_newest(self, k, buf):
        for j in range(k):
            buf[j] = torch.from_numpy(self.vecs[j])

    def import_direction(self, row):
        v = row.numpy().copy()
        self.vecs.insert(0, v)
        self.imported.append(v)


def _worker(rank, world, port, q, adaptive=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    freqs = np.linspace(8e9, 12e9, 21)
    if adaptive:       # rank 0 iterates at 4 points, rank 1 at 3: the exchange must go on for 4 rounds + 1 empty one
        eng = FakeEngine(iterating_points=4 if rank == 0 else 3)
        sh = ShardedSweep(None, freqs, rank, world, dist=dist, seed_rounds=2, engine=eng)
    else:
        eng = FakeEngine()
        sh = ShardedSweep(None, freqs, rank, world, dist=dist, seed_rounds=2, max_rounds=2, engine=eng)
    res = sh.run()
    S = sh.gather_S(res)
    q.put((rank, sh.order(), [v[0] for v in eng.imported], S[:, 0, 0].real.tolist(), sorted(res.solved), sh.rounds))
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(120)
def test_two_rank_sweep_exchanges_seed_directions_and_gathers_S():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    freqs = np.linspace(8e9, 12e9, 21)
    (r0, o0, imp0, S0, solved0, rounds0), (r1, o1, imp1, S1, solved1, rounds1) = out
    assert rounds0 == rounds1 == 2
    # contiguous blocks, each processed ends-first
    assert solved0 == list(range(0, 11)) and solved1 == list(range(11, 21))
    assert o0[:2] == [0, 10] and o1[:2] == [11, 20]
    # two seeding rounds: each rank imported the 2 ports x 2 points the other rank solved first
    exp0 = sorted(freqs[i] * 1e-9 for i in o1[:2] for _ in range(2))
    exp1 = sorted(freqs[i] * 1e-9 for i in o0[:2] for _ in range(2))
    assert np.allclose(sorted(np.real(imp0)), exp0) and np.allclose(sorted(np.real(imp1)), exp1)
    # every rank holds the S-parameters of the whole sweep
    assert np.allclose(S0, freqs * 1e-9) and np.allclose(S1, freqs * 1e-9)


@pytest.mark.timeout(120)
def test_exchange_rounds_continue_while_any_rank_adds_directions():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, True)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    freqs = np.linspace(8e9, 12e9, 21)
    (r0, o0, imp0, S0, solved0, rounds0), (r1, o1, imp1, S1, solved1, rounds1) = out
    assert rounds0 == rounds1 == 5                       # 4 rounds with new directions, then one empty round ends it
    # rank 0 imported the directions of rank 1's first 3 points, rank 1 those of rank 0's first 4 points (2 ports each)
    assert np.allclose(sorted(np.real(imp0)), sorted(freqs[i] * 1e-9 for i in o1[:3] for _ in range(2)))
    assert np.allclose(sorted(np.real(imp1)), sorted(freqs[i] * 1e-9 for i in o0[:4] for _ in range(2)))
    assert solved0 == list(range(0, 11)) and solved1 == list(range(11, 21))
    assert np.allclose(S0, freqs * 1e-9) and np.allclose(S1, freqs * 1e-9)
