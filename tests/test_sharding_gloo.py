"""The multi-GPU path is a batch split with control-plane reductions only; its host logic is covered
here with world_size = 2 on the gloo backend (CPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from intfftk_b200.sharding import reduce_report, shard_range, shard_seed


def test_shard_ranges_tile_the_batch():
    for batch in (0, 1, 7, 8, 65536, 1 << 20, 1000003):
        for world in (1, 2, 3, 4, 8):
            edges = [shard_range(batch, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == batch
            for (a0, a1), (b0, b1) in zip(edges, edges[1:]):
                assert a1 == b0 and a0 <= a1
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)
    assert shard_seed(5, 3) == 8


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import c_oracle as co
    import numpy as np
    # each rank transforms its own shard of a 10-frame job on the CPU oracle (stand-in for the GPU here)
    nfft, batch = 6, 10
    lo, hi = shard_range(batch, rank, world)
    x = co.fill_random(batch * 64 * 2, 16, 1234).reshape(batch, 64, 2)
    y = co.batch(co.generics(nfft), x[lo:hi])
    rep = reduce_report(local_ms=10.0 + rank, local_samples=(hi - lo) * 64,
                        local_checksum=co.checksum(y), dist=dist)
    q.put((rank, rep.ms, rep.samples, rep.checksum, rep.world, lo, hi))
    dist.destroy_process_group()


def test_two_rank_report_reduction_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from oracle import c_oracle as co
    x = co.fill_random(10 * 64 * 2, 16, 1234).reshape(10, 64, 2)
    y = co.batch(co.generics(6), x)
    want = (co.checksum(y[0:5]) + co.checksum(y[5:10])) & (2 ** 64 - 1)
    for rank, ms, samples, checksum, w, lo, hi in res:
        assert ms == 11.0 and samples == 640 and w == 2
        assert checksum == want
    assert [(r[5], r[6]) for r in res] == [(0, 5), (5, 10)]


def test_single_rank_report_passthrough():
    rep = reduce_report(3.5, 100, (1 << 64) + 5, None)
    assert (rep.ms, rep.samples, rep.checksum, rep.world) == (3.5, 100, 5, 1)
