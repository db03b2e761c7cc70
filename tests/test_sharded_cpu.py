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

    def __init__(self, n_ports=2, n=5, iterating_points=10 ** 9, fail_at=None):
        self.n_ports, self.n = n_ports, n
        self.vecs = []            # newest first
        self.imported = []
        self.iterating_points = iterating_points      # only the first points leave directions (the others are "accepted")
        self.solved = 0
        self.n_accepted = 0
        self.fail_at = fail_at

    def solve_point(self, f, raise_on_fail=False, out_bufs=None, keep_fields=False):
        self.solved += 1
        for p in range(self.n_ports if self.solved <= self.iterating_points else 0):
            self.vecs.insert(0, np.full(self.n, f * 1e-9 + 1j * p))
            self.n_accepted += 1
        S = np.full((self.n_ports, self.n_ports), f * 1e-9, dtype=complex)
        ok = self.fail_at is None or abs(f - self.fail_at) > 1.0
        st = [dict(freq=float(f), port=p, iters=1, relres=0.0 if ok else 0.5, converged=ok) for p in range(self.n_ports)]
        fields = {p + 1: np.full(3, f * 1e-9) for p in range(self.n_ports)} if keep_fields else {}
        return S, st, fields

    def accepted(self):
        return self.n_accepted

    def new_buffer(self, k):
        return torch.zeros((k, self.n), dtype=torch.complex128)

    def tensor(self, values, dtype=None):
        return torch.as_tensor(np.asarray(values), dtype=dtype)

    def export_newest(self, k, buf):
        for j in range(k):
            buf[j] = torch.from_numpy(self.vecs[j])

    def sync(self):
        pass

    def import_direction(self, row):
        v = row.numpy().copy()
        self.vecs.insert(0, v)
        self.n_accepted += 1
        self.imported.append(v)


def _worker(rank, world, port, q, mode="plain"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nf = 3 if mode == "tiny" else 21
    freqs = np.linspace(8e9, 12e9, nf)
    if mode == "adaptive":     # only the first 3 points a rank solves leave directions: 3 seed rounds add some, the 4th none
        eng = FakeEngine(iterating_points=3)
        sh = ShardedSweep(None, freqs, rank, world, dist=dist, engine=eng, max_seed_rounds=8)
    elif mode == "fail":
        eng = FakeEngine(fail_at=freqs[17])
        sh = ShardedSweep(None, freqs, rank, world, dist=dist, engine=eng, max_seed_rounds=2)
    else:
        eng = FakeEngine()
        sh = ShardedSweep(None, freqs, rank, world, dist=dist, engine=eng, max_seed_rounds=2)
    err = None
    fields = {}
    try:
        res = sh.run(keep_fields=True, fields_out=fields)
        S = sh.gather_S(res)
        out = (rank, sh.order(), [v[0] for v in eng.imported], S[:, 0, 0].real.tolist(), sorted(res.solved), sh.rounds,
               sorted(k[0] for k in fields), None)
    except Exception as e:           # noqa: BLE001 - the test asserts on the type name
        out = (rank, [], [], [], [], sh.rounds, [], type(e).__name__)
    q.put(out)
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run_world(world, mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, mode)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    return out


@pytest.mark.timeout(120)
def test_two_rank_sweep_seeds_globally_exchanges_and_gathers_S():
    freqs = np.linspace(8e9, 12e9, 21)
    G = hierarchical_order(21)
    (r0, o0, imp0, S0, solved0, rounds0, f0, e0), (r1, o1, imp1, S1, solved1, rounds1, f1, e1) = _run_world(2, "plain")
    assert e0 is None and e1 is None
    assert rounds0 == rounds1 == 2
    # seed rounds deal out the GLOBAL bisection order round-robin: rank 0 gets G[0], G[2], rank 1 gets G[1], G[3]
    assert o0[:2] == [G[0], G[2]] and o1[:2] == [G[1], G[3]]
    # every point is solved exactly once, the rest of each block by its owner
    assert sorted(solved0 + solved1) == list(range(21))
    seeded = set(G[:4])
    assert set(solved0) - seeded <= set(range(0, 11)) and set(solved1) - seeded <= set(range(11, 21))
    # each rank imported the 2 ports x 2 seed points of the other rank
    assert np.allclose(sorted(np.real(imp0)), sorted(freqs[i] * 1e-9 for i in o1[:2] for _ in range(2)))
    assert np.allclose(sorted(np.real(imp1)), sorted(freqs[i] * 1e-9 for i in o0[:2] for _ in range(2)))
    # every rank holds the S-parameters of the whole sweep; fields stay with the rank that solved the point
    assert np.allclose(S0, freqs * 1e-9) and np.allclose(S1, freqs * 1e-9)
    assert sorted(set(f0)) == solved0 and sorted(set(f1)) == solved1


@pytest.mark.timeout(120)
def test_seed_rounds_continue_while_any_rank_adds_directions():
    (r0, o0, imp0, S0, solved0, rounds0, _, e0), (r1, o1, imp1, S1, solved1, rounds1, _, e1) = _run_world(2, "adaptive")
    freqs = np.linspace(8e9, 12e9, 21)
    assert e0 is None and e1 is None
    assert rounds0 == rounds1 == 4                       # 3 rounds with new directions, then one empty round ends it
    assert np.allclose(sorted(np.real(imp0)), sorted(freqs[i] * 1e-9 for i in o1[:3] for _ in range(2)))
    assert np.allclose(sorted(np.real(imp1)), sorted(freqs[i] * 1e-9 for i in o0[:3] for _ in range(2)))
    assert sorted(solved0 + solved1) == list(range(21))
    assert np.allclose(S0, freqs * 1e-9) and np.allclose(S1, freqs * 1e-9)


@pytest.mark.timeout(120)
def test_unconverged_point_raises_on_every_rank_after_the_sweep():
    out = _run_world(2, "fail")
    assert [o[-1] for o in out] == ["NotConverged", "NotConverged"]


@pytest.mark.timeout(120)
def test_more_ranks_than_points():
    """world 4, 3 frequency points: rank 3 has an empty block and no seed point, and still takes part in every collective"""
    out = _run_world(4, "tiny")
    freqs = np.linspace(8e9, 12e9, 3)
    assert all(o[-1] is None for o in out)
    assert sorted(i for o in out for i in o[4]) == [0, 1, 2]
    for o in out:
        assert np.allclose(o[3], freqs * 1e-9)
