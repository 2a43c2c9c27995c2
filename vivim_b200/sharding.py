"""Clip-level data parallelism of the hot path: clips are independent sequences, so ranks own disjoint
clip ranges and no collective sits on the data path.  Only the timing
reduction (max over ranks) talks between ranks; it works on any torch.distributed backend."""
from __future__ import annotations


def clip_shard(n_clips: int, rank: int, world: int) -> range:
    """Contiguous, balanced shard of [0, n_clips) owned by `rank` (first `n_clips % world` ranks get one more)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank / world size")
    base, extra = divmod(n_clips, world)
    lo = rank * base + min(rank, extra)
    return range(lo, lo + base + (1 if rank < extra else 0))


def max_over_ranks(value: float, device="cpu") -> float:
    """All-reduce(MAX) of a per-rank scalar (e.g. the CUDA-event time of the timed region)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_throughput(units_per_rank: float, world: int, elapsed_max_s: float) -> float:
    """Whole-job throughput under weak scaling: every rank processed `units_per_rank` in `elapsed_max_s`."""
    return world * units_per_rank / elapsed_max_s
