"""N>1 path on CPU: world_size 2 over gloo.  The data path has no collective (clips are independent);
what is distributed is the clip sharding and the max-over-ranks timing reduction used by bench.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vivim_b200.sharding import aggregate_throughput, clip_shard, max_over_ranks


def test_clip_shards_partition_the_batch():
    for n, world in ((32, 8), (3, 2), (5, 8), (0, 4), (24, 1)):
        got = [i for r in range(world) for i in clip_shard(n, r, world)]
        assert got == list(range(n))
        sizes = [len(clip_shard(n, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        clip_shard(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # each rank "scans" its own clips: a CPU stand-in whose result only depends on the clip ids
        mine = clip_shard(7, rank, world)
        local = torch.tensor([float(sum(i * i for i in mine))])
        elapsed = 0.010 * (rank + 1)                       # rank 1 is the slow one
        t_max = max_over_ranks(elapsed)
        total = local.clone()
        dist.all_reduce(total)                             # test-only check that the shards cover everything
        out[rank] = (t_max, float(total.item()), aggregate_throughput(len(mine), world, t_max))
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo_max_time_and_coverage():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert set(out.keys()) == {0, 1}
    for r in range(world):
        t_max, total, _ = out[r]
        assert t_max == pytest.approx(0.020)               # max over ranks, identical on every rank
        assert total == float(sum(i * i for i in range(7)))
    # weak scaling bookkeeping: value = world * units_per_rank / max time
    assert out[0][2] == pytest.approx(2 * 4 / 0.020)


def test_single_process_reduction_is_identity():
    assert max_over_ranks(1.25) == 1.25


def test_bucket_ranges_partition_the_flat_gradient():
    from vivim_b200.graphed import bucket_ranges
    numels = [5, 100, 3, 3, 250, 1, 1, 40]
    buckets, owner = bucket_ranges(numels, 100)
    assert buckets[0][0] == 0 and buckets[-1][1] == sum(numels)
    assert all(a[1] == b[0] for a, b in zip(buckets, buckets[1:]))          # contiguous, in order
    assert all(hi - lo >= 100 for lo, hi in buckets[:-1])
    offs = [sum(numels[:i]) for i in range(len(numels))]
    for off, n, b in zip(offs, numels, owner):                                 # no tensor is split
        assert buckets[b][0] <= off and off + n <= buckets[b][1]
    assert bucket_ranges([], 10) == ([], [])


def _bucket_worker(rank, world, port, out):
    from vivim_b200.graphed import bucket_ranges
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(rank)
        numels = [7, 33, 64, 2, 90, 11]
        flat = torch.randn(sum(numels))
        whole = flat.clone()
        dist.all_reduce(whole)
        whole /= world
        buckets, _ = bucket_ranges(numels, 50)
        for lo, hi in reversed(buckets):                  # the order the backward completes them in
            part = flat[lo:hi]
            dist.all_reduce(part)
            part /= world
        out[rank] = bool(torch.equal(flat, whole)) and len(buckets) > 1
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo_bucketed_allreduce_equals_whole():
    """What TrainStepGraph captures when grad_allreduce is given: per-bucket all-reduces of views of the flat gradient
    give exactly the whole-buffer all-reduce."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_bucket_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert out[0] and out[1]
