"""World-size-2 gloo test of the multi-GPU harness logic (batch sharding + the post-run gather), on CPU."""
import importlib.util
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _shard():
    spec = importlib.util.spec_from_file_location("lbc_shard", os.path.join(ROOT, "lowbitdnn-project_b200", "shard.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["lbc_shard"] = mod
    spec.loader.exec_module(mod)
    return mod


def test_image_ranges_partition_the_batch():
    sh = _shard()
    for batch, world in [(512, 1), (512, 2), (512, 8), (7, 4), (3, 8)]:
        spans = [sh.image_range(batch, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == batch
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sh.image_range(8, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = _shard()
    first, last = sh.image_range(512, world, rank)
    # rank 1 is the slow one: the job time must be ITS time, on every rank
    st = sh.RankStats(ms_total=10.0 + 5.0 * rank, e2e_ms=2.0 + rank, checksum=1000 + rank, images=last - first,
                      bad_layers=rank, link_up=20.0 - rank, link_dn=15.0 + rank)
    job = sh.gather(st, world)
    q.put((rank, job.ms_total, job.e2e_ms, job.images, job.checksums, job.bad_layers, job.link_up, job.link_dn))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_max_over_ranks_gloo():
    world, port = 2, 29500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ms, e2e, images, sums, bad, up, dn in res:
        assert ms == 15.0 and e2e == 3.0 and images == 512 and sums == [1000, 1001]
        # a parity failure on ANY rank fails the job; the host-link figure is the slowest rank's
        assert bad == 1 and up == 19.0 and dn == 15.0
