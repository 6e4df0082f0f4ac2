"""world_size-2 gloo test (CPU) of the multi-GPU host logic: pair partition + metric reduction."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from craft_b200.sharding import pairs_for_rank, reduce_metrics


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, num_pairs, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = pairs_for_rank(num_pairs, rank, world)
    epe = sum(0.001 * (i + 1) for i in mine)          # stand-in for the per-pair EPE
    mean, total, tmax = reduce_metrics(epe, len(mine), 1.0 + rank)
    out.put((rank, mine, mean, total, tmax))
    dist.barrier()
    dist.destroy_process_group()


def test_partition_covers_every_pair_once():
    for n in (0, 1, 7, 16):
        for w in (1, 2, 3, 8):
            seen = sorted(i for r in range(w) for i in pairs_for_rank(n, r, w))
            assert seen == list(range(n))
            sizes = [len(pairs_for_rank(n, r, w)) for r in range(w)]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_reduction_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 7, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect_mean = sum(0.001 * (i + 1) for i in range(7)) / 7
    for rank, mine, mean, total, tmax in res:
        assert total == 7 and abs(mean - expect_mean) < 1e-12 and tmax == 2.0
    assert sorted(i for _, mine, *_ in res for i in mine) == list(range(7))
