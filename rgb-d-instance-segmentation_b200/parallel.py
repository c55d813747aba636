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


class GradBucketReducer:
    """Gradient all-reduce for the fine-tuning step (BASELINE configs[3]; reference: HF ``Trainer`` DDP under
    ``finetuning.py:98-113``), overlapped with the backward pass.

    Parameters are grouped into buckets in the order given (``buckets`` = list of parameter lists, e.g. one per DSAM stage:
    the backward produces dsam2's gradients first, then dsam1's, then dsam0's + DGGM's).  A post-accumulate hook on every
    parameter counts arrivals; when a bucket is complete its gradients are flattened into the bucket's buffer and
    ``all_reduce(async_op=True)`` is issued on the communication stream while autograd keeps running the earlier stages.
    ``finish()`` waits for the collectives, averages and scatters the result back into ``p.grad``.  Parameters that
    receive no gradient in a step (the ratio predictor: CM:339 consumes it through ``.item()``) must not be listed."""

    def __init__(self, buckets, average: bool = True):
        self.buckets = [list(b) for b in buckets if len(b)]
        self.average = average
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self._flat = [torch.zeros(sum(p.numel() for p in b), dtype=b[0].dtype, device=b[0].device) for b in self.buckets]
        self._pending = [0] * len(self.buckets)
        self._work = [None] * len(self.buckets)
        self._owner = {}
        self._handles = []
        for bi, b in enumerate(self.buckets):
            for p in b:
                self._owner[p] = bi
                self._handles.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self.reset()

    @property
    def n_elements(self) -> int:
        return sum(f.numel() for f in self._flat)

    def reset(self) -> None:
        self._pending = [len(b) for b in self.buckets]
        self._work = [None] * len(self.buckets)

    def _on_grad(self, p) -> None:
        bi = self._owner[p]
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            self._launch(bi)

    def _launch(self, bi: int) -> None:
        flat, off = self._flat[bi], 0
        views = []
        for p in self.buckets[bi]:
            n = p.numel()
            views.append(flat[off:off + n].view_as(p))
            off += n
        torch._foreach_copy_(views, [p.grad for p in self.buckets[bi]])
        if self.world > 1:
            self._work[bi] = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True)

    def finish(self) -> None:
        """Wait for every bucket, write the reduced (averaged) gradients back.  Buckets whose hooks did not all fire
        (a parameter without gradient this step) raise: silent partial reductions would desynchronise the ranks."""
        for bi, b in enumerate(self.buckets):
            if self._pending[bi] != 0:
                raise RuntimeError(f"bucket {bi}: {self._pending[bi]} of {len(b)} parameters received no gradient")
            if self._work[bi] is not None:
                self._work[bi].wait()
            flat, off = self._flat[bi], 0
            if self.average and self.world > 1:
                flat.div_(self.world)
            views = []
            for p in b:
                n = p.numel()
                views.append(flat[off:off + n].view_as(p))
                off += n
            torch._foreach_copy_([p.grad for p in b], views)
        self.reset()

    def remove(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []
