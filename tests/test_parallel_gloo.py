"""world_size-2 gloo test of the multi-process host logic (no GPU): frame sharding, the max-over-ranks timing
reduction and the per-rank frame bookkeeping that bench.py uses."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import rgbd_b200  # noqa: F401
from rgbd_b200 import parallel


def test_shard_ranges_partition_the_frames():
    for n in (0, 1, 7, 32, 255, 256):
        for world in (1, 2, 3, 4, 8):
            covered = []
            for r in range(world):
                b, e = parallel.shard_range(n, r, world)
                assert 0 <= b <= e <= n
                covered += list(range(b, e))
            assert covered == list(range(n))
            sizes = parallel.shard_sizes(n, world)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
    assert parallel.shard_sizes(256, 8) == [32] * 8                    # BASELINE configs[2]
    with pytest.raises(ValueError):
        parallel.shard_range(8, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    parallel.init_process_group("gloo")
    b, e = parallel.shard_range(37, rank, world)
    frames = torch.arange(b, e)
    # "hot path" stand-in: every rank works on its shard only; nothing is exchanged but time and counts
    elapsed = 0.010 * (rank + 1)
    slowest = parallel.max_over_ranks(elapsed)
    counts = parallel.gather_counts(len(frames))
    checksum = torch.tensor([float(frames.sum())], dtype=torch.float64)
    dist.all_reduce(checksum)
    # gradient buckets: reduced while "backward" is still producing the earlier ones; averaged; written back to p.grad
    torch.manual_seed(0)
    layers = [torch.nn.Linear(5, 7), torch.nn.Linear(7, 3), torch.nn.Linear(3, 2)]
    unused = torch.nn.Linear(2, 2)                      # receives no gradient: must stay out of the buckets
    red = parallel.GradBucketReducer([list(l.parameters()) for l in reversed(layers)])
    x = torch.full((4, 5), float(rank + 1))
    for step in range(2):
        for l in layers:
            l.zero_grad(set_to_none=True)
        y = x
        for l in layers:
            y = l(y)
        y.sum().backward()
        local = [p.grad.clone() for l in layers for p in l.parameters()]
        red.finish()
    reduced = [p.grad.clone() for l in layers for p in l.parameters()]
    gathered = [None, None]
    dist.all_gather_object(gathered, [g.tolist() for g in local])
    mean = [(torch.tensor(a) + torch.tensor(b)) / 2 for a, b in zip(*gathered)]
    grads_ok = all(torch.allclose(r, m, atol=1e-6) for r, m in zip(reduced, mean)) and unused.weight.grad is None
    q.put((rank, slowest, counts, float(checksum), grads_ok, red.n_elements))
    dist.destroy_process_group()


def test_two_rank_gloo_round_trip():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, slowest, counts, checksum, grads_ok, n_el in res:
        assert slowest == pytest.approx(0.020)
        assert counts == [19, 18]
        assert checksum == sum(range(37))
        assert grads_ok and n_el == 5 * 7 + 7 + 7 * 3 + 3 + 3 * 2 + 2


def test_bucket_reducer_single_process_and_incomplete_bucket():
    a, b = torch.nn.Linear(3, 3), torch.nn.Linear(3, 3)
    red = parallel.GradBucketReducer([list(a.parameters()), list(b.parameters())])
    a(torch.ones(1, 3)).sum().backward()
    with pytest.raises(RuntimeError):                 # b's bucket never filled
        red.finish()
    red.reset()
    (a(torch.ones(1, 3)).sum() + b(torch.ones(1, 3)).sum()).backward()
    before = a.weight.grad.clone()
    red.finish()                                      # world 1: gradients pass through unchanged
    assert torch.equal(a.weight.grad, before)
    red.remove()
