"""Batch sharding of the hot path across the GPUs of one box.

Every cloud is independent in forward and backward (per-sample GroupNorm, no BatchNorm,
SURVEY.md 8e), so the path shards by cloud with no data-path collective: one process per
GPU, rank r takes clouds [r*B/W, (r+1)*B/W).  Training adds exactly one exchange per step,
the sum of the weight gradients, done here as ONE flat fp32 bucket (the hot-path
parameters are ~25.5 k floats = 100 KB, so the all-reduce is latency-bound and a single
NCCL launch is the right granularity).  The reference does this with single-process
``nn.DataParallel`` (trainer_new.py:94-96).

Works with any torch.distributed backend: NCCL over NVLink on the GPU box, gloo in the
CPU tests (tests/test_parallel_cpu.py).
"""
from __future__ import annotations

import os
from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of ``global_batch`` clouds; the first ``global_batch % world``
    ranks take one extra."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(global_batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def init_from_env(backend: str | None = None) -> Tuple[int, int, int]:
    """(rank, world, local_rank) from torchrun's environment; initialises the process group
    when WORLD_SIZE > 1.  MASTER_ADDR defaults to 127.0.0.1 (single node)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


class GradBucket:
    """One flat buffer for the gradients of ``params``; ``all_reduce_mean()`` sums it over
    the ranks in a single collective and writes the averages back into ``p.grad``.

    Parameters whose ``.grad`` is None (the reference's declared-but-unused ``bn4``/``bn5``,
    M4:466-467) are skipped on every rank alike, so ranks never disagree on the layout:
    the layout is fixed by the parameter list, absent gradients are sent as zeros.
    """

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("GradBucket needs at least one trainable parameter")
        dev = self.params[0].device
        self.offsets = []
        n = 0
        for p in self.params:
            self.offsets.append(n)
            n += p.numel()
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self._zeros = None

    def all_reduce_mean(self, group=None, async_op: bool = False):
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if world == 1:
            return None
        # pack with one kernel (absent gradients go in as zeros), average inside the collective where the backend can
        if self._zeros is None:
            self._zeros = [torch.zeros(p.numel(), dtype=torch.float32, device=self.flat.device) for p in self.params]
        srcs = [z if p.grad is None else p.grad.reshape(-1) for p, z in zip(self.params, self._zeros)]
        torch.cat(srcs, out=self.flat)
        in_collective = dist.get_backend(group) == "nccl"
        op = dist.ReduceOp.AVG if in_collective else dist.ReduceOp.SUM
        work = dist.all_reduce(self.flat, op=op, group=group, async_op=async_op)
        div = 1 if in_collective else world
        if async_op:
            return _Pending(self, work, div)
        self._scatter_back(div)
        return None

    def _scatter_back(self, world: int):
        if world != 1:
            self.flat.div_(world)
        dsts, views = [], []
        for p, o in zip(self.params, self.offsets):
            if p.grad is not None:
                dsts.append(p.grad)
                views.append(self.flat[o:o + p.numel()].view_as(p.grad))
        if dsts:
            torch._foreach_copy_(dsts, views)          # one multi-tensor kernel


class _Pending:
    def __init__(self, bucket, work, world):
        self.bucket, self.work, self.world = bucket, work, world

    def wait(self):
        self.work.wait()
        self.bucket._scatter_back(self.world)
