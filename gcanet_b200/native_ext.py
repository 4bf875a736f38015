"""Loader of the compiled PyTorch C++ extension (``gcanet_b200/csrc_ext/torch_ext.cpp``).

The reference reaches its native kernels through pybind11 modules built by torch's cpp_extension: ``knn.knn(ref, query, k)``
(models/KNN_CUDA/knn_cuda/csrc/cuda/knn.cpp:59-61) and ``_ext.group_points`` / ``_ext.group_points_grad``
(PN2 _ext-src/src/bindings.cpp:17-18).  ``load()`` returns the module that plays both roles here -- same function names,
argument order, dtypes, shapes and error behaviour, every function one call into ``libgcanet_b200.so`` -- and, as a side
effect of importing it, registers ``torch.ops.gcanet_b200_native.{knn, knn_graph, group_points, group_points_grad}``
(CUDA dispatch key only).  This module adds what belongs on the Python side of such an extension: fake (meta)
implementations for FakeTensor tracing / ``torch.compile`` and the autograd formula of ``group_points``.

    from gcanet_b200 import native_ext
    ext = native_ext.load()
    dist, ind = ext.knn(ref, query, k)                       # [dim, Nr], [dim, Nq] -> [k, Nq] each, 1-based like the reference
    out = torch.ops.gcanet_b200_native.group_points(features, idx32)      # differentiable

The ctypes front end (``gcanet_b200.functional``) stays the default path of the package; both end in the same kernels.
Built ahead of time by ``python -m gcanet_b200.build``; a missing file raises (no JIT, no fallback).
"""
from __future__ import annotations

import importlib.util
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
EXT_PATH = os.path.join(_HERE, "lib", "gcanet_b200_ext.so")
_NAME = "gcanet_b200_ext"
_NS = "gcanet_b200_native"

_ext = None
_lock = threading.Lock()


def _kout(k1: int, k2: int) -> int:
    step = k2 // k1
    return (k2 + step - 1) // step


def _register_python_side() -> None:
    @torch.library.register_fake(f"{_NS}::knn")
    def _(ref, query, k, index_base=1):
        shape = (ref.shape[0], k, query.shape[-1]) if ref.dim() == 3 else (k, query.shape[-1])
        return query.new_empty(shape), query.new_empty(shape, dtype=torch.int64)

    @torch.library.register_fake(f"{_NS}::knn_graph")
    def _(x, k1, k2, metric=0):
        return x.new_empty((x.shape[0], x.shape[2], _kout(k1, k2)), dtype=torch.int64)

    @torch.library.register_fake(f"{_NS}::group_points")
    def _(points, idx):
        return points.new_empty((points.shape[0], points.shape[1], idx.shape[1], idx.shape[2]))

    @torch.library.register_fake(f"{_NS}::group_points_grad")
    def _(grad_out, idx, n):
        return grad_out.new_empty((grad_out.shape[0], grad_out.shape[1], n))

    # GroupingOperation.backward (PN2/pointnet2_utils.py:218-237): grad wrt the features only
    def _setup(ctx, inputs, output):
        points, idx = inputs
        ctx.save_for_backward(idx)
        ctx.n = points.shape[2]

    def _backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return torch.ops.gcanet_b200_native.group_points_grad(grad_out.contiguous(), idx, ctx.n), None

    torch.library.register_autograd(f"{_NS}::group_points", _backward, setup_context=_setup)


def load():
    """Imports gcanet_b200/lib/gcanet_b200_ext.so once and returns the module.  Raises when it has not been built."""
    global _ext
    if _ext is None:
        with _lock:
            if _ext is None:
                if not os.path.exists(EXT_PATH):
                    raise RuntimeError(f"{EXT_PATH} is missing: build it with `python -m gcanet_b200.build`")
                spec = importlib.util.spec_from_file_location(_NAME, EXT_PATH)
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                if mod.abi_version() != 2:
                    raise RuntimeError("gcanet_b200_ext.so was built against another ABI version of libgcanet_b200.so")
                _register_python_side()
                _ext = mod
    return _ext
