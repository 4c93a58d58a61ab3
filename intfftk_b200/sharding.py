"""Batch split across GPUs: frames are independent, so a job of `batch` frames is cut
into contiguous shards, one per rank, with NO exchange step on the data path.  The only collectives
are control-plane reductions of the report (elapsed time = max over ranks, sample counts and shard
checksums = sum), done with torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from dataclasses import dataclass


def shard_range(batch: int, rank: int, world: int) -> tuple[int, int]:
    """Frames [lo, hi) owned by `rank`; shards differ by at most one frame and tile the batch."""
    if not (0 <= rank < world) or batch < 0:
        raise ValueError("bad shard request")
    return batch * rank // world, batch * (rank + 1) // world


def shard_seed(seed: int, rank: int) -> int:
    """Every rank draws its stimulus from its own counter stream of the shared seed."""
    return (seed + rank) & (2 ** 64 - 1)


@dataclass
class JobReport:
    ms: float            # max over ranks
    samples: int         # sum over ranks
    checksum: int        # sum over ranks modulo 2^64 (a checksum of shard checksums)
    world: int


def reduce_report(local_ms: float, local_samples: int, local_checksum: int, dist=None, device=None) -> JobReport:
    """Combine per-rank results. `dist` is torch.distributed (initialised) or None for one rank."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return JobReport(float(local_ms), int(local_samples), int(local_checksum) & (2 ** 64 - 1), 1)
    import torch
    t = torch.tensor([local_ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # 64-bit modular sum done on 32-bit halves in int64 (no unsigned all-reduce in torch)
    lo, hi = local_checksum & 0xFFFFFFFF, (local_checksum >> 32) & 0xFFFFFFFF
    s = torch.tensor([local_samples, lo, hi], dtype=torch.int64, device=device)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    total = (int(s[1]) + (int(s[2]) << 32)) & (2 ** 64 - 1)
    return JobReport(float(t[0]), int(s[0]), total, dist.get_world_size())
