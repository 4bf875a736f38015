"""``torch.ops.gcanet_b200.*``: the hot-path entry points registered with the PyTorch dispatcher (``torch.library``).

The reference binds its native code as pybind11 modules built by torch's cpp_extension (``knn.knn(ref, query, k)``,
models/KNN_CUDA/knn_cuda/csrc/cuda/knn.cpp:59-61; ``_ext.group_points`` / ``_ext.group_points_grad``,
PN2 _ext-src/src/bindings.cpp:17-18).  Here the kernels live in a PyTorch-free C-ABI library; this module is the thin
PyTorch-side registration on top of it: each operator has a CUDA implementation (a ctypes call into
``libgcanet_b200.so``), a fake (meta) implementation, so FakeTensor tracing / ``torch.compile`` see shapes and dtypes
without running a kernel, and, where the reference op is differentiable, a registered backward that is itself an
operator.  Nothing here adds a code path: the functions of ``gcanet_b200.functional`` remain the implementations.

    import gcanet_b200.torch_ops                      # registers the operators
    idx = torch.ops.gcanet_b200.knn_graph(x, 50, 50, 0, True)
    out_nc, out_cn = gcanet_b200.torch_ops.edgeconv(x_nc, idx32, weight, gamma, beta, C)      # differentiable
"""
from __future__ import annotations

import ctypes as _ct
from typing import Tuple

import torch
from torch import Tensor

from . import _cabi
from . import functional as G
from ._cabi import EdgeConvDesc, call, ptr, stream, workspace

_NS = "gcanet_b200"


def _kout(k1: int, k2: int) -> int:
    step = k2 // k1
    return (k2 + step - 1) // step


# ------------------------------------------------------------------ kNN graph (M4:30-90)
@torch.library.custom_op(f"{_NS}::knn_graph", mutates_args=(), device_types="cuda")
def knn_graph(x: Tensor, k1: int, k2: int, metric: int, ordered: bool) -> Tensor:
    """x [B, C, N] fp32 -> idx [B, N, kout] int64; metric 0 = L2 (``knn``), 1 = points x normals (``knn_points_normals``)."""
    return G.knn_graph(x, k1, k2, metric, want64=True, want32=False, ordered=ordered)[0]


@knn_graph.register_fake
def _(x, k1, k2, metric, ordered):
    return x.new_empty((x.shape[0], x.shape[2], _kout(k1, k2)), dtype=torch.int64)


# ------------------------------------------------------------------ KNN_CUDA (knn.cpp:23-56)
@torch.library.custom_op(f"{_NS}::knn_cuda", mutates_args=(), device_types="cuda")
def knn_cuda(ref: Tensor, query: Tensor, k: int) -> Tuple[Tensor, Tensor]:
    """ref [B, dim, Nr], query [B, dim, Nq] -> (dist [B, k, Nq] Euclidean, idx [B, k, Nq] int64, 0-based)."""
    d, i = G.knn_cuda(ref, query, k, index_base=0)
    return d, i


@knn_cuda.register_fake
def _(ref, query, k):
    B, nq = ref.shape[0], query.shape[2]
    return ref.new_empty((B, k, nq)), ref.new_empty((B, k, nq), dtype=torch.int64)


# ------------------------------------------------------------------ grouping (bindings.cpp:17-18)
@torch.library.custom_op(f"{_NS}::group_points", mutates_args=(), device_types="cuda")
def group_points(features: Tensor, idx: Tensor) -> Tensor:
    """features [B, C, N] fp32, idx [B, npoint, nsample] int32 -> [B, C, npoint, nsample]."""
    return G.GroupingOperation.forward(_NoCtx(), features, idx)


@group_points.register_fake
def _(features, idx):
    return features.new_empty((features.shape[0], features.shape[1], idx.shape[1], idx.shape[2]))


@torch.library.custom_op(f"{_NS}::group_points_grad", mutates_args=(), device_types="cuda")
def group_points_grad(grad_out: Tensor, idx: Tensor, n: int) -> Tensor:
    g = grad_out.contiguous()
    B, C, npnt, ns = g.shape
    with torch.cuda.device(g.device):
        gp = torch.empty((B, C, n), dtype=torch.float32, device=g.device)
        call("gcanet_group_points_grad", B, C, n, npnt, ns, ptr(g), ptr(idx), ptr(gp), stream())
    return gp


@group_points_grad.register_fake
def _(grad_out, idx, n):
    return grad_out.new_empty((grad_out.shape[0], grad_out.shape[1], n))


class _NoCtx:
    """Stand-in for the autograd context when a Function's forward is reused as a plain implementation."""

    def save_for_backward(self, *a):
        pass


def _group_points_setup(ctx, inputs, output):
    features, idx = inputs
    ctx.save_for_backward(idx)
    ctx.n = features.shape[2]


def _group_points_backward(ctx, grad_out):
    (idx,) = ctx.saved_tensors
    return torch.ops.gcanet_b200.group_points_grad(grad_out, idx, ctx.n), None


group_points.register_autograd(_group_points_backward, setup_context=_group_points_setup)


# ------------------------------------------------------------------ fused EdgeConv block (M4:469-481, 494-505)
def _desc(x_nc, idx32, weight, C, groups, eps, slope):
    B, N, ldx = x_nc.shape
    return EdgeConvDesc(B, N, C, ldx, weight.shape[0], idx32.shape[2], groups, eps, slope, 0)


@torch.library.custom_op(f"{_NS}::edgeconv_forward", mutates_args=(), device_types="cuda")
def edgeconv_forward(x_nc: Tensor, idx32: Tensor, weight: Tensor, gamma: Tensor, beta: Tensor, C: int, groups: int,
                     eps: float, slope: float) -> Tuple[Tensor, Tensor, Tensor]:
    """-> (out_nc [B, N, Cout], out_cn [B, Cout, N], saved: opaque bytes the backward needs)."""
    desc = _desc(x_nc, idx32, weight, C, groups, eps, slope)
    L = _cabi.lib()
    B, N, _ = x_nc.shape
    Cout = weight.shape[0]
    with torch.cuda.device(x_nc.device):
        saved_bytes = L.gcanet_edgeconv_saved_bytes(_ct.byref(desc))
        if saved_bytes == 0:
            raise RuntimeError("gcanet_b200 edgeconv: " + L.gcanet_last_error().decode())
        saved = torch.empty(saved_bytes + 256, dtype=torch.uint8, device=x_nc.device)      # owned by autograd, not the pool
        base = saved[(-saved.data_ptr()) % 256:][:saved_bytes]                              # 256-byte aligned view
        ws = workspace(L.gcanet_edgeconv_workspace_bytes(_ct.byref(desc)), x_nc.device)
        out_nc = torch.empty((B, N, Cout), dtype=torch.float32, device=x_nc.device)
        out_cn = torch.empty((B, Cout, N), dtype=torch.float32, device=x_nc.device)
        call("gcanet_edgeconv_forward", _ct.byref(desc), ptr(x_nc), ptr(idx32), ptr(weight), ptr(gamma), ptr(beta),
             ptr(out_nc), ptr(out_cn), ptr(base), ptr(ws), ws.numel(), stream())
    return out_nc, out_cn, saved


@edgeconv_forward.register_fake
def _(x_nc, idx32, weight, gamma, beta, C, groups, eps, slope):
    B, N, _ = x_nc.shape
    Cout = weight.shape[0]
    # size of the saved state: a host-side query of the library (no GPU needed)
    nbytes = _cabi.lib().gcanet_edgeconv_saved_bytes(_ct.byref(_desc(x_nc, idx32, weight, C, groups, eps, slope))) + 256
    return x_nc.new_empty((B, N, Cout)), x_nc.new_empty((B, Cout, N)), x_nc.new_empty((nbytes,), dtype=torch.uint8)


@torch.library.custom_op(f"{_NS}::edgeconv_backward", mutates_args=(), device_types="cuda")
def edgeconv_backward(x_nc: Tensor, idx32: Tensor, weight: Tensor, gamma: Tensor, beta: Tensor, grad_nc: Tensor,
                      saved: Tensor, C: int, groups: int, eps: float, slope: float) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    desc = _desc(x_nc, idx32, weight, C, groups, eps, slope)
    L = _cabi.lib()
    with torch.cuda.device(x_nc.device):
        base = saved[(-saved.data_ptr()) % 256:]
        g = grad_nc.contiguous()
        gx, gw, gg, gb = torch.empty_like(x_nc), torch.empty_like(weight), torch.empty_like(gamma), torch.empty_like(beta)
        ws = workspace(L.gcanet_edgeconv_workspace_bytes(_ct.byref(desc)), x_nc.device)
        call("gcanet_edgeconv_backward", _ct.byref(desc), ptr(x_nc), ptr(idx32), ptr(weight), ptr(gamma), ptr(beta), ptr(g),
             ptr(base), ptr(gx), ptr(gw), ptr(gg), ptr(gb), ptr(ws), ws.numel(), stream())
    return gx, gw, gg, gb


@edgeconv_backward.register_fake
def _(x_nc, idx32, weight, gamma, beta, grad_nc, saved, C, groups, eps, slope):
    return torch.empty_like(x_nc), torch.empty_like(weight), torch.empty_like(gamma), torch.empty_like(beta)


def _edgeconv_setup(ctx, inputs, output):
    x_nc, idx32, weight, gamma, beta, C, groups, eps, slope = inputs
    ctx.save_for_backward(x_nc, idx32, weight, gamma, beta, output[2])
    ctx.args = (C, groups, eps, slope)


def _edgeconv_backward(ctx, g_nc, g_cn, g_saved):
    x_nc, idx32, weight, gamma, beta, saved = ctx.saved_tensors
    C, groups, eps, slope = ctx.args
    g = g_nc
    if g_cn is not None:
        g = g_cn.transpose(1, 2) if g is None else g + g_cn.transpose(1, 2)
    if g is None:
        g = torch.zeros((x_nc.shape[0], x_nc.shape[1], weight.shape[0]), dtype=torch.float32, device=x_nc.device)
    gx, gw, gg, gb = torch.ops.gcanet_b200.edgeconv_backward(x_nc, idx32, weight, gamma, beta, g.contiguous(), saved, C,
                                                            groups, eps, slope)
    return gx, None, gw, gg, gb, None, None, None, None


edgeconv_forward.register_autograd(_edgeconv_backward, setup_context=_edgeconv_setup)


def edgeconv(x_nc, idx32, weight, gamma, beta, C, groups=2, eps=1e-5, slope=0.2):
    """Dispatcher-registered form of ``functional.edgeconv`` (fp32 storage): (out_nc, out_cn), differentiable."""
    w2 = weight.reshape(weight.shape[0], -1).contiguous()
    out_nc, out_cn, _ = torch.ops.gcanet_b200.edgeconv_forward(x_nc, idx32, w2, gamma.contiguous(), beta.contiguous(), int(C),
                                                               int(groups), float(eps), float(slope))
    return out_nc, out_cn
