"""Multi-GPU plumbing: independent RGB-D streams are partitioned across ranks; no data-path collective.

The only cross-rank traffic is the benchmark's barrier and the max-over-ranks of the elapsed time (torch.distributed:
NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations


def shard_streams(n_streams: int, world: int, rank: int) -> list[int]:
    """stream s -> rank s mod world (one or more sequences per GPU)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    return [s for s in range(n_streams) if s % world == rank]


def stream_seed(rank: int, local_index: int, streams_per_rank: int) -> int:
    """global stream id of the local_index-th stream of a rank when every rank owns streams_per_rank streams (weak scaling)."""
    return rank * streams_per_rank + local_index


def max_over_ranks(x: float, dist=None, device=None) -> float:
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    import torch

    t = torch.tensor([float(x)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_rate(units_per_rank: int, world: int, max_seconds: float) -> float:
    """whole-job throughput: units all ranks processed / slowest rank's time."""
    return units_per_rank * world / max_seconds
