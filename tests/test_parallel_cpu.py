"""CPU suite, part 3: the N > 1 path (batch sharding + one flat gradient all-reduce) on the
gloo backend with world_size 2 -- the host logic bench.py runs over NCCL on the GPU box."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gcanet_b200.parallel import GradBucket, shard_range


def test_shard_range_is_a_balanced_partition():
    for gb, world in [(128, 8), (128, 2), (16, 1), (10, 4), (3, 8)]:
        spans = [shard_range(gb, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == gb
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            assert a1 == b0
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    assert shard_range(128, 3, 8) == (48, 64)          # BASELINE config 4: 16 clouds per GPU at 8 GPUs
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)                      # identical replicas
        lin = torch.nn.Linear(4, 3)
        unused = torch.nn.Parameter(torch.zeros(5))        # like bn4/bn5: never gets a gradient
        params = list(lin.parameters()) + [unused]
        bucket = GradBucket(params)
        # each rank owns its shard of a global batch of 6 samples
        g = torch.Generator().manual_seed(1)
        data = torch.randn(6, 4, generator=g)
        lo, hi = shard_range(6, rank, world)
        (lin(data[lo:hi]).sum() / 6.0 * world).backward()   # so that the rank-mean equals the global-batch gradient
        bucket.all_reduce_mean()
        assert unused.grad is None
        out[rank] = [p.grad.clone() for p in lin.parameters()]
        # async flavour gives the same result
        lin.zero_grad()
        (lin(data[lo:hi]).sum() / 6.0 * world).backward()
        bucket.all_reduce_mean(async_op=True).wait()
        for a, b in zip(out[rank], [p.grad for p in lin.parameters()]):
            assert torch.allclose(a, b)
        # a gradient that exists on rank 0 only: every rank must end up with the average (replicas diverged)
        lonely = torch.nn.Parameter(torch.zeros(3))
        b2 = GradBucket([lonely])
        if rank == 0:
            lonely.grad = torch.full((3,), 4.0)
        b2.all_reduce_mean()
        assert lonely.grad is not None and torch.allclose(lonely.grad, torch.full((3,), 4.0 / world))
    finally:
        dist.destroy_process_group()


def test_grad_bucket_allreduce_matches_global_batch_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    torch.manual_seed(0)
    lin = torch.nn.Linear(4, 3)
    g = torch.Generator().manual_seed(1)
    data = torch.randn(6, 4, generator=g)
    (lin(data).sum() / 6.0).backward()
    want = [p.grad for p in lin.parameters()]
    for r in range(world):
        for a, b in zip(out[r], want):
            assert torch.allclose(a, b, atol=1e-6), (r, a, b)


def test_single_process_bucket_is_a_no_op():
    lin = torch.nn.Linear(2, 2)
    lin(torch.ones(1, 2)).sum().backward()
    before = [p.grad.clone() for p in lin.parameters()]
    assert GradBucket(lin.parameters()).all_reduce_mean() is None
    GradBucket(lin.parameters()).all_reduce_mean(async_op=True).wait()       # a no-op handle, not None
    for a, b in zip(before, [p.grad for p in lin.parameters()]):
        assert torch.equal(a, b)
