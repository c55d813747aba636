"""Batch data-parallelism for the hot path: every image is independent, so frames are
sharded across ranks (one process per GPU) and NO collective touches the data path.  The only communication is
the max-over-ranks of the measured time (and an optional gather of results to rank 0)."""
from __future__ import annotations

import os
from typing import List, Tuple

import torch
import torch.distributed as dist


def env_rank_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of ``n_items`` frames owned by ``rank``; sizes differ by at most one."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    per, rem = divmod(n_items, world)
    begin = rank * per + min(rank, rem)
    return begin, begin + per + (1 if rank < rem else 0)


def shard_sizes(n_items: int, world: int) -> List[int]:
    return [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)]


def init_process_group(backend: str, device=None) -> None:
    """Initialise torch.distributed from the torchrun environment (MASTER_ADDR should be 127.0.0.1 on one node)."""
    if dist.is_initialized():
        return
    kwargs = {}
    if backend == "nccl" and device is not None:
        kwargs["device_id"] = device
    dist.init_process_group(backend, **kwargs)


def max_over_ranks(value: float, device="cpu") -> float:
    """Whole-job time = the slowest rank's device time."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_counts(n_local: int, device="cpu") -> List[int]:
    """Frames processed per rank, gathered everywhere (bookkeeping for the whole-job throughput)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [n_local]
    t = torch.tensor([n_local], dtype=torch.int64, device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [int(x.item()) for x in out]
