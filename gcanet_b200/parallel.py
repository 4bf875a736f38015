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
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)      # binds the communicator to this GPU (barrier() needs no guess)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


class GradBucket:
    """One flat buffer for the gradients of ``params``; ``all_reduce_mean()`` sums it over
    the ranks in a single collective and writes the averages back into ``p.grad``.

    The layout is fixed by the parameter list, so ranks never disagree on it.  A parameter whose
    ``.grad`` is None (the reference's declared-but-unused ``bn4``/``bn5``, M4:466-467) is sent as
    zeros, followed by one presence flag per parameter; after the collective a parameter that got a
    gradient on ANY rank has the average written back on EVERY rank (``p.grad`` is created where it was
    missing), so replicas cannot drift apart when a gradient is absent on some ranks only.  Reading the
    flags costs one small device-to-host copy per step; ``assume_uniform=True`` skips it for callers
    that know every rank produces the same set of gradients (bench.py's hot-path parameters).

    The average is a mean of per-rank means: it equals the global-batch gradient when every rank holds
    the same number of clouds (what ``shard_range`` gives when world divides the global batch);
    otherwise scale each rank's loss by ``local_batch * world / global_batch`` before backward.
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], assume_uniform: bool = False):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("GradBucket needs at least one trainable parameter")
        dev = self.params[0].device
        self.offsets = []
        n = 0
        for p in self.params:
            self.offsets.append(n)
            n += p.numel()
        self.numel = n
        self.assume_uniform = assume_uniform
        # gradients, then one presence flag per parameter
        self.flat = torch.zeros(n + len(self.params), dtype=torch.float32, device=dev)
        self._zeros = None
        self._has_sinks = False
        self._flag_cache = {}

    def attach_sinks(self):
        """Let the fused backward kernels write parameter gradients directly into this bucket (functional._grad_sinks):
        autograd then adopts the bucket slice as ``p.grad`` and ``all_reduce_mean`` finds nothing to pack or to copy
        back.  Safe with any caller: a gradient that did not land in its slice (another code path, accumulation into an
        existing ``.grad``) is packed and scattered back the ordinary way.  Callers must reset ``p.grad = None`` before
        each backward -- accumulating into the slice a kernel is about to overwrite would double-count."""
        from . import functional as G
        for p, o in zip(self.params, self.offsets):
            G._grad_sinks[p.data_ptr()] = [self.flat[o:o + p.numel()], False]
        self._has_sinks = True
        return self

    def release_sinks(self):
        """Start of a new step: every sink may be handed out again (``all_reduce_mean`` does this itself)."""
        from . import functional as G
        for p in self.params:
            e = G._grad_sinks.get(p.data_ptr())
            if e is not None:
                e[1] = False

    def detach_sinks(self):
        from . import functional as G
        for p in self.params:
            G._grad_sinks.pop(p.data_ptr(), None)

    def _in_place(self, p, o):
        return p.grad is not None and p.grad.data_ptr() == self.flat.data_ptr() + 4 * o and p.grad.is_contiguous()

    def _flags(self, present):
        key = tuple(present)
        t = self._flag_cache.get(key)
        if t is None:
            t = torch.tensor([1.0 if f else 0.0 for f in present], dtype=torch.float32, device=self.flat.device)
            self._flag_cache[key] = t
        return t

    def all_reduce_mean(self, group=None, async_op: bool = False):
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if self._has_sinks:
            self.release_sinks()
        if world == 1:
            return _Done() if async_op else None
        # pack with one kernel (absent gradients go in as zeros), average inside the collective where the backend can
        if self._zeros is None:
            self._zeros = [torch.zeros(p.numel(), dtype=torch.float32, device=self.flat.device) for p in self.params]
        present = [p.grad is not None for p in self.params]
        if all(self._in_place(p, o) for p, o in zip(self.params, self.offsets)):
            # every gradient was written into its slice by the backward kernels (attach_sinks): only the flags are set
            self.flat[self.numel:].copy_(self._flags(present))
        else:
            srcs = [z if p.grad is None else p.grad.reshape(-1) for p, z in zip(self.params, self._zeros)]
            torch.cat(srcs + [self._flags(present)], out=self.flat)
        in_collective = dist.get_backend(group) == "nccl"
        op = dist.ReduceOp.AVG if in_collective else dist.ReduceOp.SUM
        work = dist.all_reduce(self.flat, op=op, group=group, async_op=async_op)
        div = 1 if in_collective else world
        if async_op:
            return _Pending(self, work, div, present)
        self._scatter_back(div, present)
        return None

    def _scatter_back(self, world: int, present):
        if world != 1:
            self.flat.div_(world)
        if self.assume_uniform:
            anywhere = present
        else:
            anywhere = (self.flat[self.numel:] > 0).tolist()          # one small D2H copy
        dsts, views = [], []
        for p, o, here, some in zip(self.params, self.offsets, present, anywhere):
            if not some:
                continue
            view = self.flat[o:o + p.numel()].view_as(p)
            if here and self._in_place(p, o):
                continue                                # p.grad IS the slice: the collective already averaged it
            if here:
                dsts.append(p.grad)
                views.append(view.view_as(p.grad))
            else:
                p.grad = view.clone()             # absent here, present elsewhere: take the average all the same
        if dsts:
            torch._foreach_copy_(dsts, views)          # one multi-tensor kernel


class _Pending:
    def __init__(self, bucket, work, world, present):
        self.bucket, self.work, self.world, self.present = bucket, work, world, present

    def wait(self):
        self.work.wait()
        self.bucket._scatter_back(self.world, self.present)


class _Done:
    """What ``all_reduce_mean(async_op=True)`` returns in a single-process run: nothing to wait for."""

    def wait(self):
        return None
