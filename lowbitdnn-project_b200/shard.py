"""Batch sharding and result gathering for the multi-GPU harness.

The convolution path has no collective: every rank owns `batch_per_gpu` whole images and a replica of the weights.
The only communication is an all-gather of per-rank timings and output checksums AFTER the timed region
(NCCL on GPUs; the same code runs under gloo in the CPU tests)."""
from __future__ import annotations

from dataclasses import dataclass


def image_range(global_batch: int, world: int, rank: int) -> tuple[int, int]:
    """[first, last) image indices of `rank` when `global_batch` images are split as evenly as possible."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} / world {world}")
    base, extra = divmod(global_batch, world)
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


@dataclass
class RankStats:
    ms_total: float      # device time of the K timed steps on this rank
    e2e_ms: float        # per-step end-to-end (host buffers) time on this rank
    checksum: int        # crc32 of this rank's output of the last layer
    images: int          # images this rank processed per step
    bad_layers: int = 0  # layer outputs that differ from the oracle on this rank (0 = parity)
    link_up: float = 0.0   # pinned host->device GB/s of this rank with all ranks copying at once
    link_dn: float = 0.0   # device->host GB/s, same condition


@dataclass
class JobStats:
    ms_total: float      # max over ranks
    e2e_ms: float        # max over ranks
    images: int          # sum over ranks
    checksums: list
    bad_layers: int = 0  # sum over ranks
    link_up: float = 0.0   # min over ranks
    link_dn: float = 0.0


def gather(stats: RankStats, world: int, device=None) -> JobStats:
    """All ranks call this; every rank gets the job-level numbers (max time, summed images, every checksum)."""
    if world == 1:
        return JobStats(stats.ms_total, stats.e2e_ms, stats.images, [stats.checksum], stats.bad_layers, stats.link_up, stats.link_dn)
    import torch
    import torch.distributed as dist
    t = torch.tensor([stats.ms_total, stats.e2e_ms, float(stats.checksum), float(stats.images), float(stats.bad_layers),
                      stats.link_up, stats.link_dn], dtype=torch.float64,
                     device=device if device is not None else "cpu")
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    rows = [o.cpu().tolist() for o in out]
    return JobStats(max(r[0] for r in rows), max(r[1] for r in rows), int(sum(r[3] for r in rows)), [int(r[2]) for r in rows],
                    int(sum(r[4] for r in rows)), min(r[5] for r in rows), min(r[6] for r in rows))
