"""GPU: the frequency-sharded sweep against the single-rank sweep and the reference's S-parameters.

Two ranks run ShardedSweep on the wg_medium fixture refined to 41 points: seed rounds over the global bisection order,
exchange of reduced-basis directions through emb_recycle_export / emb_recycle_import, fill of the contiguous blocks,
all-reduce of the S-parameters.  With two GPUs the ranks use NCCL and their own devices; on a one-GPU box both ranks share
cuda:0 and the exchange is staged through the host (gloo), which exercises the same C-ABI calls."""
import os
import socket

import numpy as np
import pytest

from tests.util import load_golden, golden_bcs, db_deg_close

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, backend, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = rank if backend == "nccl" else 0
    torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    from emerge_b200.sweep import FrequencySweep
    from emerge_b200.distributed import ShardedSweep
    g, t = load_golden("wg_medium")
    freqs = np.linspace(8e9, 12e9, 41)
    sw = FrequencySweep(t, g["er"], g["ur"], golden_bcs(g, t), device=dev)
    sw.f_ref = 10e9
    sw.setup()
    sh = ShardedSweep(sw, freqs, rank, world, dist=dist, device=dev)
    res = sh.run()
    S = sh.gather_S(res)
    q.put((rank, S, sorted(res.solved), sh.rounds, sh.exchanged, res.max_relres, sw.ctx.recycle_info()["n"]))
    sw.ctx.close()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_sharded_sweep_equals_single_rank_and_reference():
    import torch
    import torch.multiprocessing as mp
    from emerge_b200.sweep import FrequencySweep
    world = 2
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, backend, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted((q.get(timeout=500) for _ in range(world)), key=lambda o: o[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g, t = load_golden("wg_medium")
    freqs = np.linspace(8e9, 12e9, 41)
    sw = FrequencySweep(t, g["er"], g["ur"], golden_bcs(g, t))
    single = sw.run(list(freqs))
    sw.ctx.close()
    (r0, S0, solved0, rounds0, ex0, rel0, n0), (r1, S1, solved1, rounds1, ex1, rel1, n1) = out
    assert sorted(solved0 + solved1) == list(range(41)) and not set(solved0) & set(solved1)
    assert rounds0 == rounds1 >= 1 and ex0 > 0 and ex1 > 0          # directions did cross the rank boundary
    assert rel0 <= 1e-8 and rel1 <= 1e-8
    assert np.array_equal(S0, S1)                                    # every rank holds the same gathered S
    assert db_deg_close(S0, single.S), np.abs(S0 - single.S).max()   # sharded == single rank within 1e-3 dB / 0.1 deg
    idx = [int(np.argmin(np.abs(freqs - f))) for f in g["freqs"]]
    assert db_deg_close(S0[idx], g["S"])                             # == the reference at the fixture's frequencies
