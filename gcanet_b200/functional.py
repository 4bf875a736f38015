"""Reference-signature front end of the B200 kNN-graph + EdgeConv path.

Every public function keeps the name, argument order, defaults, return arity, dtypes
and tensor layouts of the reference function it replaces (cited per function; paths are
relative to the reference root, M4 = models/dgcnn-hais-concat-direct-4.py), so a call
site can switch by changing its import.  All compute goes through the C-ABI of
libgcanet_b200.so; there is no CPU or eager fallback -- CPU tensors raise.
"""
from __future__ import annotations

import ctypes as _ct

import torch

from . import _cabi
from ._cabi import EdgeConvDesc, GlobalFeatureDesc, GroupNormDesc, NormalEdgeDesc, OffsetDesc, call, ptr, require_cuda, stream, workspace

METRIC_L2 = 0
METRIC_POINTS_NORMALS = 1
EDGE_DIFF_CENTER = 0
EDGE_NORMAL_ANGLE = 1



# Optional per-call device timing (bench.py): CUDA events on the launching stream around
# each C-ABI compute call.  Off by default; costs two event records per call when on.
_timing = None


def enable_kernel_timing(flag: bool = True) -> None:
    global _timing
    _timing = [] if flag else None


def kernel_timings_ms() -> dict:
    """{tag: [ms, ...]} for the calls recorded since the last enable; call after a synchronize."""
    out = {}
    for tag, a, b in (_timing or []):
        out.setdefault(tag, []).append(a.elapsed_time(b))
    return out


# With `scan_probe=True` the per-call timing also asks the library (gcanet_knn_probe_arm / _read) for the time of the
# tensor-core distance-scan kernel(s) inside each kNN call -- the dominant kernel of the step -- and records it under
# the call's tag + ":scan_kernel".  The read waits for the device after every kNN call: breakdown passes only.
_scan_probe = False
_scan_ms: dict = {}


def enable_scan_probe(flag: bool = True) -> None:
    global _scan_probe
    _scan_probe = bool(flag)
    _scan_ms.clear()


def scan_kernel_timings_ms() -> dict:
    """{kNN call tag: [ms of its scan kernel(s), ...]} recorded while the scan probe was on."""
    return {k: list(v) for k, v in _scan_ms.items()}


class _timed:
    def __init__(self, tag):
        self.tag = tag

    def __enter__(self):
        if _timing is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if _timing is not None:
            self.b.record()
            _timing.append((self.tag, self.a, self.b))
        return False


# Gradient sinks (gcanet_b200.parallel.GradBucket.attach_sinks): parameter storage address -> [slice of a flat all-reduce
# bucket shaped like the parameter, handed-out flag].  A fused backward that finds an unused sink writes the parameter's
# gradient straight into the bucket, so nothing has to be packed before the collective or copied back after it.  A sink
# is handed out once per step (the bucket clears the flags when it reduces): a second backward call for the same
# parameter in the same step -- weight sharing, or a batch split over two streams -- gets an ordinary buffer, which
# autograd then ADDS into the slice that already is ``p.grad``.
_grad_sinks: dict = {}


def _grad_buffer(param_like: torch.Tensor) -> torch.Tensor:
    entry = _grad_sinks.get(param_like.data_ptr())
    if entry is not None and not entry[1] and entry[0].numel() == param_like.numel() and entry[0].device == param_like.device:
        entry[1] = True
        return entry[0].view(param_like.shape)
    return torch.empty_like(param_like)


def _as_f32_contig(x: torch.Tensor, name: str) -> torch.Tensor:
    require_cuda(x, name, contiguous=False)
    if x.dtype != torch.float32:
        x = x.float()
    return x.contiguous()


# ----------------------------------------------------------------------------------
# kNN graph (torch path)
# ----------------------------------------------------------------------------------
KNN_FLAG_BRUTE_FORCE = 0x100
KNN_FLAG_UNORDERED = 0x200
KNN_FLAG_NO_PRUNE = 0x400


def knn_graph(x: torch.Tensor, k1: int, k2: int, metric: int = METRIC_L2, want64: bool = True,
              want32: bool = False, tensor_cores: bool = True, brute_force: bool = False, ordered: bool = True,
              prune: bool = True):
    """x [B, C, N] -> (idx64 or None, idx32 or None), each [B, N, kout].  ``brute_force=True`` (or
    the older ``tensor_cores=False``) forces the plain CUDA-core scan where an accelerated path
    (tcgen05 pruning for C = 64/128, spatial pruning for xyz clouds) would apply; ``prune=False`` makes
    the tensor-core path scan every key tile instead of skipping by bounding box -- for A/B tests."""
    if brute_force or not tensor_cores:
        metric = metric | KNN_FLAG_BRUTE_FORCE
    if not prune:
        metric = metric | KNN_FLAG_NO_PRUNE
    if not ordered and int(k1) == int(k2):
        metric = metric | KNN_FLAG_UNORDERED      # same exact neighbour set, order unspecified
    x = _as_f32_contig(x.detach(), "x")
    if x.dim() != 3:
        raise RuntimeError(f"x must be [B, C, N] (got {tuple(x.shape)})")
    B, C, N = x.shape
    k1, k2 = int(k1), int(k2)
    L = _cabi.lib()
    kout = L.gcanet_knn_graph_columns(k1, k2)
    if kout <= 0:
        raise RuntimeError(f"need 1 <= k1 <= k2 (k1={k1}, k2={k2})")
    with torch.cuda.device(x.device):
        i64 = torch.empty((B, N, kout), dtype=torch.int64, device=x.device) if want64 else None
        i32 = torch.empty((B, N, kout), dtype=torch.int32, device=x.device) if want32 else None
        ws_bytes = L.gcanet_knn_graph_workspace_bytes(B, C, N, k2, metric)
        ws = workspace(ws_bytes, x.device)
        tag = f"knn_graph[C={C},metric={metric & 0xff}]"
        probe = _scan_probe and not torch.cuda.is_current_stream_capturing()
        if probe:
            call("gcanet_knn_probe_arm", 1)
        with _timed(tag):
            call("gcanet_knn_graph", ptr(x), B, C, N, k1, k2, metric, ptr(i64), ptr(i32), ptr(ws), ws.numel(), stream())
        if probe:
            ms = _ct.c_float(0.0)
            if L.gcanet_knn_probe_read(_ct.byref(ms)) == 0:       # non-zero: this call took a path without a tensor-core scan
                _scan_ms.setdefault(tag, []).append(float(ms.value))
            call("gcanet_knn_probe_arm", 0)
    return i64, i32


def knn(x, k1, k2=None):
    """Replaces ``knn(x, k1, k2)`` (M4:30-47) and, called with two arguments, ``knn(x, k)``
    of models/splinenet.py:9-22.  x [B, C, N] -> idx [B, N, k1] int64, nearest first, the
    point itself included, reference dilation columns ``arange(0, k2, k2 // k1)``."""
    if k2 is None:
        k2 = k1
    return knn_graph(x, k1, k2, METRIC_L2)[0]


def knn_points_normals(x, k1, k2):
    """Replaces ``knn_points_normals`` (M4:50-90); x is [B, 6, N] (xyz + unit normals)."""
    return knn_graph(x, k1, k2, METRIC_POINTS_NORMALS)[0]


# ----------------------------------------------------------------------------------
# materialised graph features
# ----------------------------------------------------------------------------------
class _GraphFeature(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, idx, variant):
        B, C, N = x.shape
        k = idx.shape[2]
        L = _cabi.lib()
        F = L.gcanet_graph_feature_channels(C, variant)
        with torch.cuda.device(x.device):
            out = torch.empty((B, N, k, F), dtype=torch.float32, device=x.device)
            ws = workspace(L.gcanet_graph_feature_workspace_bytes(B, C, N, k, variant), x.device)
            call("gcanet_graph_feature", ptr(x), ptr(idx), ptr(out), B, C, N, k, variant, ptr(ws), ws.numel(), stream())
        ctx.save_for_backward(x, idx)
        ctx.variant = variant
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, idx = ctx.saved_tensors
        B, C, N = x.shape
        k = idx.shape[2]
        L = _cabi.lib()
        g = grad_out.contiguous().float()
        with torch.cuda.device(x.device):
            gx = torch.empty_like(x)
            ws = workspace(L.gcanet_graph_feature_grad_workspace_bytes(B, C, N, k, ctx.variant), x.device)
            call("gcanet_graph_feature_grad", ptr(g), ptr(x), ptr(idx), ptr(gx), B, C, N, k, ctx.variant,
                 ptr(ws), ws.numel(), stream())
        return gx, None, None


def _graph_feature(x, k1, k2, idx, metric, variant):
    B, N = x.size(0), x.size(2)
    x = _as_f32_contig(x.reshape(B, -1, N), "x")
    if idx is None:
        idx = knn_graph(x, k1, k2, metric)[0]
    else:
        require_cuda(idx, "idx", contiguous=False)
        idx = idx.to(torch.int64).contiguous()
        if idx.dim() != 3 or idx.shape[0] != B or idx.shape[1] != N:
            raise RuntimeError(f"idx must be [B, N, k] (got {tuple(idx.shape)})")
        if idx.shape[2] != k1:
            # the reference's view(batch, num_points, k1, dims) fails the same way (M4:120)
            raise RuntimeError(f"idx has {idx.shape[2]} columns but k1={k1}")
    out = _GraphFeature.apply(x, idx, variant)
    return out.permute(0, 3, 1, 2)          # the reference returns this permuted view (M4:123)


def get_graph_feature(x, k1=20, k2=20, idx=None):
    """Replaces ``get_graph_feature`` (M4:93-124): [B, C, N] -> [B, 2C, N, k1], a permuted view
    of a contiguous [B, N, k1, 2C] buffer holding (x_j - x_i, x_i); differentiable in x."""
    return _graph_feature(x, k1, k2, idx, METRIC_L2, EDGE_DIFF_CENTER)


def get_graph_feature_with_normals(x, k1=20, k2=20, idx=None):
    """Replaces ``get_graph_feature_with_normals`` (M4:127-161): neighbours under the
    points x normals metric, same (x_j - x_i, x_i) feature on all 6 channels."""
    return _graph_feature(x, k1, k2, idx, METRIC_POINTS_NORMALS, EDGE_DIFF_CENTER)


def get_graph_feature_with_normals_g(x, k1=20, k2=20, idx=None):
    """Replaces ``get_graph_feature_with_normals_g`` (M4:164-205): [B, 6, N] -> [B, 7, N, k1] =
    (clamp(n_i . n_j, -0.99, 0.99), n_j - n_i, n_i)."""
    return _graph_feature(x, k1, k2, idx, METRIC_POINTS_NORMALS, EDGE_NORMAL_ANGLE)


def splinenet_knn(x, k):
    """``knn(x, k)`` of models/splinenet.py:9-22."""
    return knn(x, k, k)


def splinenet_get_graph_feature(x, k=20, idx=None):
    """``get_graph_feature(x, k=20, idx=None)`` of models/splinenet.py:25-53."""
    return get_graph_feature(x, k1=k, k2=k, idx=idx)


# ----------------------------------------------------------------------------------
# KNN_CUDA path
# ----------------------------------------------------------------------------------
def knn_cuda(ref: torch.Tensor, query: torch.Tensor, k: int, index_base: int = 0):
    """Batched core of the KNN_CUDA replacement.  ref [B, dim, Nr], query [B, dim, Nq] ->
    (dist [B, k, Nq] fp32 Euclidean, ind [B, k, Nq] int64)."""
    require_cuda(ref, "ref", torch.float32)
    require_cuda(query, "query", torch.float32)
    if ref.dim() != 3 or query.dim() != 3 or ref.shape[0] != query.shape[0] or ref.shape[1] != query.shape[1]:
        raise RuntimeError(f"ref.shape={tuple(ref.shape)} != query.shape={tuple(query.shape)}")
    B, dim, nr = ref.shape
    nq = query.shape[2]
    L = _cabi.lib()
    with torch.cuda.device(ref.device):
        dist = torch.empty((B, k, nq), dtype=torch.float32, device=ref.device)
        ind = torch.empty((B, k, nq), dtype=torch.int64, device=ref.device)
        ws = workspace(L.gcanet_knn_cuda_workspace_bytes(B, dim, nr, nq, k), ref.device)
        call("gcanet_knn_cuda", ptr(ref), nr, ptr(query), nq, dim, int(k), B, index_base, ptr(dist), ptr(ind),
             ptr(ws), ws.numel(), stream())
    return dist, ind


def knn_cuda_pair(ref, query, k):
    """Replaces ``knn(ref, query, k)`` of models/KNN_CUDA/knn_cuda/__init__.py:41-44:
    ref [dim, Nr], query [dim, Nq] -> (dist [k, Nq], ind [k, Nq] int64, 0-based)."""
    d, i = knn_cuda(ref.contiguous()[None], query.contiguous()[None], k, index_base=0)
    return d[0], i[0]


class KNN(torch.nn.Module):
    """Replaces ``KNN(k, transpose_mode=False)`` (models/KNN_CUDA/knn_cuda/__init__.py:54-74).
    transpose_mode=False: ref [B, dim, Nr], query [B, dim, Nq] -> D, I [B, k, Nq];
    transpose_mode=True : ref [B, Nr, dim], query [B, Nq, dim] -> D, I [B, Nq, k].
    The reference loops over the batch in Python; here the batch is one launch."""

    def __init__(self, k, transpose_mode=False):
        super().__init__()
        self.k = k
        self._t = transpose_mode

    def forward(self, ref, query):
        assert ref.size(0) == query.size(0), "ref.shape={} != query.shape={}".format(ref.shape, query.shape)
        with torch.no_grad():
            if self._t:
                ref, query = ref.transpose(1, 2), query.transpose(1, 2)
            D, I = knn_cuda(ref.float().contiguous(), query.float().contiguous(), self.k, index_base=0)
            if self._t:
                D, I = D.transpose(1, 2).contiguous(), I.transpose(1, 2).contiguous()
        return D, I


# ----------------------------------------------------------------------------------
# PN2 grouping
# ----------------------------------------------------------------------------------
class GroupingOperation(torch.autograd.Function):
    """Replaces ``GroupingOperation`` (PN2 pointnet2_utils.py:194-240): features [B, C, N] fp32
    contiguous, idx [B, npoint, nsample] int32 contiguous -> [B, C, npoint, nsample].  Like the
    reference's C++ checks (group_points.cpp:13-16) wrong dtype / layout / device raises."""

    @staticmethod
    def forward(ctx, features, idx):
        require_cuda(features, "features", torch.float32)
        require_cuda(idx, "idx", torch.int32)
        B, C, N = features.shape
        _, npnt, ns = idx.shape
        with torch.cuda.device(features.device):
            out = torch.empty((B, C, npnt, ns), dtype=torch.float32, device=features.device)
            call("gcanet_group_points", B, C, N, npnt, ns, ptr(features), ptr(idx), ptr(out), stream())
        ctx.save_for_backward(idx)
        ctx.n = N
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        g = grad_out.contiguous()
        B, C, npnt, ns = g.shape
        with torch.cuda.device(g.device):
            gp = torch.empty((B, C, ctx.n), dtype=torch.float32, device=g.device)
            call("gcanet_group_points_grad", B, C, ctx.n, npnt, ns, ptr(g), ptr(idx), ptr(gp), stream())
        return gp, None


grouping_operation = GroupingOperation.apply


def knn_point(group_size, point_cloud, query_cloud):
    """Replaces ``knn_point`` (models/search_knn.py:11-14): (dist, idx), both [B, k, Nq]."""
    return KNN(k=group_size, transpose_mode=False)(point_cloud, query_cloud)


def group_points(group_size, point_cloud, query_cloud, point_features=None):
    """Replaces ``group_points`` (models/search_knn.py:23-39): returns
    (grouped_points [B, 3, Nq, k], grouped_features [B, F, Nq, k] or None, idx [B, Nq, k] int32)."""
    _, idx = knn_point(group_size, point_cloud, query_cloud)
    idx = idx.permute(0, 2, 1).type(torch.int32).contiguous()
    grouped_points = grouping_operation(point_cloud.contiguous(), idx)
    grouped_features = None if point_features is None else grouping_operation(point_features.contiguous(), idx)
    return grouped_points, grouped_features, idx


# ----------------------------------------------------------------------------------
# fused EdgeConv block
# ----------------------------------------------------------------------------------
def to_point_major(x_cn: torch.Tensor, ld: int | None = None, add: torch.Tensor | None = None) -> torch.Tensor:
    """[B, C, N] -> [B, N, ld] (ld = C rounded up to a multiple of 4 by default, zero padded);
    with ``add`` [B, N, ld] the result is transpose(x) + add in the same pass."""
    require_cuda(x_cn, "x", torch.float32)
    B, C, N = x_cn.shape
    ld = ld or (C + 3) // 4 * 4
    with torch.cuda.device(x_cn.device):
        out = torch.empty((B, N, ld), dtype=torch.float32, device=x_cn.device)
        if add is None:
            call("gcanet_cn_to_nc", ptr(x_cn), ptr(out), B, C, N, ld, stream())
        else:
            require_cuda(add, "add", torch.float32)
            if tuple(add.shape) != (B, N, ld):
                raise RuntimeError(f"add must be [B, N, ld] = {(B, N, ld)} (got {tuple(add.shape)})")
            call("gcanet_cn_to_nc_add", ptr(x_cn), ptr(add), ptr(out), B, C, N, ld, stream())
    return out


def to_channel_major(x_nc: torch.Tensor, C: int | None = None) -> torch.Tensor:
    require_cuda(x_nc, "x", torch.float32)
    B, N, ld = x_nc.shape
    C = C or ld
    with torch.cuda.device(x_nc.device):
        out = torch.empty((B, C, N), dtype=torch.float32, device=x_nc.device)
        call("gcanet_nc_to_cn", ptr(x_nc), ptr(out), B, C, N, ld, stream())
    return out


class _ToPointMajor(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_cn, ld):
        ctx.C = x_cn.shape[1]
        return to_point_major(x_cn, ld)

    @staticmethod
    def backward(ctx, g):
        return to_channel_major(g.contiguous(), ctx.C), None


class _EdgeConv(torch.autograd.Function):
    """out_nc, out_cn = EdgeConv(x_nc, idx32, weight [Cout, 2C], gamma, beta)."""

    @staticmethod
    def forward(ctx, x_nc, idx32, weight, gamma, beta, C, groups, eps, slope, want_cn, storage_bf16):
        require_cuda(x_nc, "x_nc", torch.float32)
        require_cuda(idx32, "idx", torch.int32)
        weight = weight.contiguous()
        gamma, beta = gamma.contiguous(), beta.contiguous()
        for t, nme in ((weight, "weight"), (gamma, "gamma"), (beta, "beta")):
            require_cuda(t, nme, torch.float32)
        B, N, ldx = x_nc.shape
        k = idx32.shape[2]
        Cout = weight.shape[0]
        if weight.shape[1] != 2 * C:
            raise RuntimeError(f"weight must be [Cout, {2 * C}] (got {tuple(weight.shape)})")
        if tuple(idx32.shape[:2]) != (B, N):
            raise RuntimeError(f"idx must be [B, N, k] (got {tuple(idx32.shape)})")
        desc = EdgeConvDesc(B, N, C, ldx, Cout, k, groups, eps, slope, 1 if storage_bf16 else 0)
        L = _cabi.lib()
        with torch.cuda.device(x_nc.device):
            saved_bytes = L.gcanet_edgeconv_saved_bytes(_ct.byref(desc))
            ws_bytes = L.gcanet_edgeconv_workspace_bytes(_ct.byref(desc))
            if saved_bytes == 0:
                raise RuntimeError("gcanet_b200 edgeconv: " + L.gcanet_last_error().decode())
            saved = workspace(saved_bytes, x_nc.device)
            ws = workspace(ws_bytes, x_nc.device)
            out_nc = torch.empty((B, N, Cout), dtype=torch.float32, device=x_nc.device)
            out_cn = torch.empty((B, Cout, N), dtype=torch.float32, device=x_nc.device) if want_cn else None
            with _timed(f"edgeconv_fwd[C={C},Cout={Cout}]"):
                call("gcanet_edgeconv_forward", _ct.byref(desc), ptr(x_nc), ptr(idx32), ptr(weight), ptr(gamma),
                     ptr(beta), ptr(out_nc), ptr(out_cn), ptr(saved), ptr(ws), ws.numel(), stream())
        ctx.save_for_backward(x_nc, idx32, weight, gamma, beta, saved)
        ctx.desc = desc
        ctx.want_cn = want_cn
        ctx.set_materialize_grads(False)      # an unused layout of the output hands back None, not a zero tensor to add
        return out_nc, out_cn

    @staticmethod
    def backward(ctx, g_nc, g_cn):
        x_nc, idx32, weight, gamma, beta, saved = ctx.saved_tensors
        desc = ctx.desc
        L = _cabi.lib()
        with torch.cuda.device(x_nc.device):
            g = g_nc.contiguous() if g_nc is not None else None
            if g_cn is not None:
                g = to_point_major(g_cn.contiguous(), desc.Cout, add=g)
            if g is None:
                g = torch.zeros((desc.B, desc.N, desc.Cout), dtype=torch.float32, device=x_nc.device)
            need_x = ctx.needs_input_grad[0]
            gx = torch.empty_like(x_nc) if need_x else None
            gw, gg, gb = _grad_buffer(weight), _grad_buffer(gamma), _grad_buffer(beta)
            ws = workspace(L.gcanet_edgeconv_workspace_bytes(_ct.byref(desc)), x_nc.device)
            with _timed(f"edgeconv_bwd[C={desc.C},Cout={desc.Cout}]"):
                call("gcanet_edgeconv_backward", _ct.byref(desc), ptr(x_nc), ptr(idx32), ptr(weight), ptr(gamma),
                     ptr(beta), ptr(g), ptr(saved), ptr(gx), ptr(gw), ptr(gg), ptr(gb), ptr(ws), ws.numel(), stream())
        return gx, None, gw, gg, gb, None, None, None, None, None, None


def edgeconv(x_nc, idx32, weight, gamma, beta, C, groups=2, eps=1e-5, slope=0.2, want_cn=True, storage="fp32"):
    """Fused replacement of ``get_graph_feature(x, idx=idx) -> Conv2d(2C, Cout, 1, bias=False) ->
    GroupNorm(groups, Cout) -> LeakyReLU(slope) -> max over k`` (M4:469-481, M4:494-505).

    x_nc [B, N, ld] point-major (``to_point_major``), idx32 [B, N, k] int32, weight [Cout, 2C]
    (or the conv's [Cout, 2C, 1, 1]).  Returns (out_nc [B, N, Cout], out_cn [B, Cout, N] or None);
    differentiable in x_nc, weight, gamma, beta.  ``storage="bf16"`` keeps the projected operand [P|Q] in bf16 (the
    "bf16 activations" mode of BASELINE configs[1]: half the gather traffic and half the saved bytes, outputs within bf16
    rounding of the pre-norm activation); the default "fp32" is the parity mode."""
    if storage not in ("fp32", "bf16"):
        raise ValueError("storage must be 'fp32' or 'bf16'")
    w2 = weight.reshape(weight.shape[0], -1)
    return _EdgeConv.apply(x_nc, idx32, w2, gamma, beta, int(C), int(groups), float(eps), float(slope), bool(want_cn),
                           storage == "bf16")


class _NormalEdgeConv(torch.autograd.Function):
    """out_cn = max_k LReLU(GN(conv_normal(edge feature of normals)))  -- parameter gradients only."""

    @staticmethod
    def forward(ctx, x_nc, idx32, weight, gamma, beta, groups, eps, slope):
        require_cuda(x_nc, "x_nc", torch.float32)
        require_cuda(idx32, "idx", torch.int32)
        weight, gamma, beta = weight.contiguous(), gamma.contiguous(), beta.contiguous()
        B, N, ldx = x_nc.shape
        k = idx32.shape[2]
        Cout = weight.shape[0]
        if weight.shape[1] != 7:
            raise RuntimeError(f"conv_normal weight must be [Cout, 7] (got {tuple(weight.shape)})")
        desc = NormalEdgeDesc(B, N, ldx, Cout, k, groups, eps, slope)
        L = _cabi.lib()
        with torch.cuda.device(x_nc.device):
            saved_bytes = L.gcanet_normal_edgeconv_saved_bytes(_ct.byref(desc))
            if saved_bytes == 0:
                raise RuntimeError("gcanet_b200 normal_edgeconv: " + L.gcanet_last_error().decode())
            saved = workspace(saved_bytes, x_nc.device)
            ws = workspace(L.gcanet_normal_edgeconv_workspace_bytes(_ct.byref(desc)), x_nc.device)
            out_nc = torch.empty((B, N, Cout), dtype=torch.float32, device=x_nc.device)
            out_cn = torch.empty((B, Cout, N), dtype=torch.float32, device=x_nc.device)
            with _timed(f"normal_edgeconv_fwd[Cout={Cout}]"):
                call("gcanet_normal_edgeconv_forward", _ct.byref(desc), ptr(x_nc), ptr(idx32), ptr(weight), ptr(gamma),
                     ptr(beta), ptr(out_nc), ptr(out_cn), ptr(saved), ptr(ws), ws.numel(), stream())
        ctx.save_for_backward(x_nc, idx32, weight, gamma, beta, saved)
        ctx.desc = desc
        return out_cn

    @staticmethod
    def backward(ctx, g_cn):
        x_nc, idx32, weight, gamma, beta, saved = ctx.saved_tensors
        desc = ctx.desc
        L = _cabi.lib()
        with torch.cuda.device(x_nc.device):
            g = to_point_major(g_cn.contiguous(), desc.Cout)
            gw, gg, gb = torch.empty_like(weight), torch.empty_like(gamma), torch.empty_like(beta)
            ws = workspace(L.gcanet_normal_edgeconv_workspace_bytes(_ct.byref(desc)), x_nc.device)
            with _timed(f"normal_edgeconv_bwd[Cout={desc.Cout}]"):
                call("gcanet_normal_edgeconv_backward", _ct.byref(desc), ptr(x_nc), ptr(idx32), ptr(weight), ptr(gamma),
                     ptr(beta), ptr(g), ptr(saved), ptr(gw), ptr(gg), ptr(gb), ptr(ws), ws.numel(), stream())
        return None, None, gw, gg, gb, None, None, None


def normal_edgeconv(points, idx32, weight, gamma, beta, groups=2, eps=1e-5, slope=0.2):
    """Fused replacement of ``conv_normal(get_graph_feature_with_normals_g(points, idx=idx)).max(-1)[0]``
    (M4:584-587, M4:691-693).  points [B, 6, N] (xyz + normals, data -- no gradient flows to it),
    idx32 [B, N, k] int32, weight [64, 7] or [64, 7, 1, 1].  Returns [B, Cout, N]; differentiable in
    weight, gamma, beta."""
    x_nc = to_point_major(points.detach().float().contiguous(), 8)
    return _NormalEdgeConv.apply(x_nc, idx32, weight.reshape(weight.shape[0], -1), gamma, beta, int(groups),
                                 float(eps), float(slope))


# ----------------------------------------------------------------------------------
# encoder tail: Conv1d(256 -> 1024) + GroupNorm + ReLU + max over the points, fused
# ----------------------------------------------------------------------------------
class _GlobalFeature(torch.autograd.Function):
    """x4 [B, Cout] = max_n relu(GroupNorm(Conv1d(x)))  from point-major x_nc [B, N, K]."""

    @staticmethod
    def forward(ctx, x_nc, weight, bias, gamma, beta, groups, eps):
        require_cuda(x_nc, "x_nc", torch.float32)
        weight, gamma, beta = weight.contiguous(), gamma.contiguous(), beta.contiguous()
        bias = None if bias is None else bias.contiguous()
        for t, nme in ((weight, "weight"), (gamma, "gamma"), (beta, "beta")):
            require_cuda(t, nme, torch.float32)
        B, N, K = x_nc.shape
        Cout = weight.shape[0]
        if weight.shape[1] != K:
            raise RuntimeError(f"weight must be [Cout, {K}] (got {tuple(weight.shape)})")
        desc = GlobalFeatureDesc(B, N, K, Cout, groups, eps)
        L = _cabi.lib()
        with torch.cuda.device(x_nc.device):
            saved_bytes = L.gcanet_global_feature_saved_bytes(_ct.byref(desc))
            if saved_bytes == 0:
                raise RuntimeError("gcanet_b200 global_feature: " + L.gcanet_last_error().decode())
            saved = workspace(saved_bytes, x_nc.device)
            ws = workspace(L.gcanet_global_feature_workspace_bytes(_ct.byref(desc)), x_nc.device)
            out = torch.empty((B, Cout), dtype=torch.float32, device=x_nc.device)
            with _timed(f"global_feature_fwd[K={K},Cout={Cout}]"):
                call("gcanet_global_feature_forward", _ct.byref(desc), ptr(x_nc), ptr(weight), ptr(bias), ptr(gamma), ptr(beta),
                     ptr(out), ptr(saved), ptr(ws), ws.numel(), stream())
        ctx.save_for_backward(x_nc, weight, gamma, beta, saved)
        ctx.bias = bias
        ctx.desc = desc
        return out

    @staticmethod
    def backward(ctx, g):
        x_nc, weight, gamma, beta, saved = ctx.saved_tensors
        bias, desc = ctx.bias, ctx.desc
        L = _cabi.lib()
        with torch.cuda.device(x_nc.device):
            g = g.contiguous().float()
            gx = torch.empty_like(x_nc) if ctx.needs_input_grad[0] else None
            gw, gg, gb = torch.empty_like(weight), torch.empty_like(gamma), torch.empty_like(beta)
            gbias = None if bias is None else torch.empty_like(bias)
            ws = workspace(L.gcanet_global_feature_workspace_bytes(_ct.byref(desc)), x_nc.device)
            with _timed(f"global_feature_bwd[K={desc.K},Cout={desc.Cout}]"):
                call("gcanet_global_feature_backward", _ct.byref(desc), ptr(x_nc), ptr(weight), ptr(bias), ptr(gamma), ptr(beta),
                     ptr(g), ptr(saved), ptr(gx), ptr(gw), ptr(gbias), ptr(gg), ptr(gb), ptr(ws), ws.numel(), stream())
        return gx, gw, gbias, gg, gb, None, None


def global_feature(x_nc, weight, bias, gamma, beta, groups=8, eps=1e-5):
    """Fused replacement of ``relu(bnmlp1(mlp1(x_features))).max(dim=2)[0]`` (M4:507-510): x_nc [B, N, 256] point-major
    concatenation of x1 | x2 | x3, weight [1024, 256] (or the Conv1d's [1024, 256, 1]), bias [1024] or None.
    Returns x4 [B, Cout]; differentiable in x_nc, weight, bias, gamma, beta.  The [B, Cout, N] activation is never formed."""
    return _GlobalFeature.apply(x_nc, weight.reshape(weight.shape[0], -1), bias, gamma, beta, int(groups), float(eps))


# ----------------------------------------------------------------------------------
# GroupNorm + ReLU of the per-point heads (M4:644-713), channel-major
# ----------------------------------------------------------------------------------
class _GroupNormAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, groups, eps, act):
        require_cuda(x, "x", torch.float32)
        if x.dim() != 3:
            raise RuntimeError(f"group_norm_act: x must be [B, C, N] (got {tuple(x.shape)})")
        x = x.contiguous()
        gamma, beta = gamma.contiguous(), beta.contiguous()
        B, C, N = x.shape
        if gamma.numel() != C or beta.numel() != C:
            raise RuntimeError(f"group_norm_act: weight / bias must have C = {C} elements")
        desc = GroupNormDesc(B, C, N, groups, eps, act)
        L = _cabi.lib()
        with torch.cuda.device(x.device):
            ws_bytes = L.gcanet_group_norm_workspace_bytes(_ct.byref(desc))
            if ws_bytes == 0:
                raise RuntimeError(f"group_norm_act: groups = {groups} must divide C = {C}")
            ws = workspace(ws_bytes, x.device)
            y = torch.empty_like(x)
            stats = torch.empty((B, groups, 2), dtype=torch.float32, device=x.device)
            with _timed("group_norm_fwd"):
                call("gcanet_group_norm_forward", _ct.byref(desc), ptr(x), ptr(gamma), ptr(beta), ptr(y), ptr(stats), ptr(ws),
                     ws.numel(), stream())
        ctx.save_for_backward(x, gamma, beta, stats)
        ctx.desc = desc
        return y

    @staticmethod
    def backward(ctx, gy):
        x, gamma, beta, stats = ctx.saved_tensors
        desc = ctx.desc
        L = _cabi.lib()
        with torch.cuda.device(x.device):
            gy = gy.contiguous().float()
            gx, gg, gb = torch.empty_like(x), torch.empty_like(gamma), torch.empty_like(beta)
            ws = workspace(L.gcanet_group_norm_workspace_bytes(_ct.byref(desc)), x.device)
            with _timed("group_norm_bwd"):
                call("gcanet_group_norm_backward", _ct.byref(desc), ptr(x), ptr(gamma), ptr(beta), ptr(stats), ptr(gy), ptr(gx),
                     ptr(gg), ptr(gb), ptr(ws), ws.numel(), stream())
        return gx, gg, gb, None, None, None


def group_norm_relu(x, weight, bias, groups, eps=1e-5):
    """``F.relu(GroupNorm(groups, C)(x))`` on channel-major x [B, C, N] (M4:644-645, 650, 661, 698, 713): one statistics
    pass spread over the whole GPU, one normalise + ReLU pass; backward in three passes.  Differentiable in x, weight, bias."""
    return _GroupNormAct.apply(x, weight, bias, int(groups), float(eps), 1)


def group_norm(x, weight, bias, groups, eps=1e-5):
    """``GroupNorm(groups, C)(x)`` on channel-major x [B, C, N], same kernels without the activation."""
    return _GroupNormAct.apply(x, weight, bias, int(groups), float(eps), 0)


# ----------------------------------------------------------------------------------
# offset-prediction block (OFFSET_PRED_MODULE + KPAM + cos_dist, M4:326-452), fused
# ----------------------------------------------------------------------------------
class _OffsetPred(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, feature, inst, key_index, conv_w, gamma, beta, w1, w2, off_w, off_b, k, groups, eps, slope):
        for t, nme in ((points, "points"), (feature, "feature"), (inst, "instance_feature")):
            require_cuda(t, nme, torch.float32)
        require_cuda(key_index, "key_index", torch.int32)
        params = [t.contiguous() for t in (conv_w, gamma, beta, w1, w2, off_w, off_b)]
        B, N, _ = points.shape
        if tuple(points.shape) != (B, N, 3) or tuple(feature.shape) != (B, N, 128) or inst.shape[:2] != (B, N):
            raise RuntimeError(f"offset_pred: points [B,N,3], feature [B,N,128], instance_feature [B,N,E] expected "
                               f"(got {tuple(points.shape)}, {tuple(feature.shape)}, {tuple(inst.shape)})")
        desc = OffsetDesc(B, N, key_index.numel(), k, inst.shape[2], groups, eps, slope)
        L = _cabi.lib()
        with torch.cuda.device(points.device):
            saved_bytes = L.gcanet_offset_pred_saved_bytes(_ct.byref(desc))
            if saved_bytes == 0:
                raise RuntimeError("gcanet_b200 offset_pred: " + L.gcanet_last_error().decode())
            saved = workspace(saved_bytes, points.device)
            ws = workspace(L.gcanet_offset_pred_workspace_bytes(_ct.byref(desc)), points.device)
            out = torch.empty((B, 3, N), dtype=torch.float32, device=points.device)
            with _timed("offset_pred_fwd"):
                call("gcanet_offset_pred_forward", _ct.byref(desc), ptr(points), ptr(feature), ptr(inst), ptr(key_index),
                     *[ptr(t) for t in params], ptr(out), ptr(saved), ptr(ws), ws.numel(), stream())
        ctx.save_for_backward(points, feature, inst, key_index, saved, *params)
        ctx.desc = desc
        return out

    @staticmethod
    def backward(ctx, g):
        points, feature, inst, key_index, saved, *params = ctx.saved_tensors
        desc = ctx.desc
        L = _cabi.lib()
        with torch.cuda.device(points.device):
            g = g.contiguous().float()
            gf, gi = torch.empty_like(feature), torch.empty_like(inst)
            gp = [torch.empty_like(t) for t in params]
            ws = workspace(L.gcanet_offset_pred_workspace_bytes(_ct.byref(desc)), points.device)
            with _timed("offset_pred_bwd"):
                call("gcanet_offset_pred_backward", _ct.byref(desc), ptr(points), ptr(feature), ptr(inst), ptr(key_index),
                     *[ptr(t) for t in params], ptr(g), ptr(saved), ptr(gf), ptr(gi), *[ptr(t) for t in gp], ptr(ws),
                     ws.numel(), stream())
        return (None, gf, gi, None, *gp, None, None, None, None)


def offset_pred(points, feature, instance_feature, key_index, conv_w, gamma, beta, att_w1, att_w2, off_w, off_b, k=30,
                groups=2, eps=1e-5, slope=0.2):
    """Fused ``OFFSET_PRED_MODULE.forward`` (M4:398-452): points [B, N, 3] (data), feature [B, N, 128], instance_feature
    [B, N, E], key_index [S] int32 -> offsets [B, 3, N]; differentiable in feature, instance_feature and all parameters."""
    return _OffsetPred.apply(points.contiguous(), feature.contiguous(), instance_feature.contiguous(), key_index,
                             conv_w.reshape(conv_w.shape[0], -1), gamma, beta, att_w1.reshape(att_w1.shape[0], -1),
                             att_w2.reshape(att_w2.shape[0], -1), off_w.reshape(off_w.shape[0], -1), off_b, int(k), int(groups),
                             float(eps), float(slope))


# ----------------------------------------------------------------------------------
# dense affinity + gated ball query (front end of the proposal grouping, M4:210-233, M4:1215-1233)
# ----------------------------------------------------------------------------------
def compute_batch_adjacency_matrix(batch_point_clouds, radius=0, dist_state=True, sigma=1.0):
    """Replaces ``compute_batch_adjacency_matrix`` (M4:210-233) for the way the model calls it (``dist_state=True`` on one
    [M, C] cloud, M4:1215-1217): exp(-(d / d_max)^2 / (2 sigma^2)) with a zero diagonal, as a dense [M, M] (or [1, M, M])
    matrix.  The fused ``affinity_ball_query`` answers the same question without building it."""
    if not dist_state:
        raise NotImplementedError("dist_state=False (thresholded adjacency) is not used by the model and not built")
    x = batch_point_clouds
    squeeze = x.dim() == 2
    if x.dim() == 3:
        if x.shape[0] != 1:
            raise NotImplementedError("one cloud per call (the reference normalises over the whole batch tensor; its call sites pass one)")
        x = x[0]
    require_cuda(x, "batch_point_clouds", contiguous=False)
    x = x.float().contiguous()
    M, C = x.shape
    L = _cabi.lib()
    with torch.cuda.device(x.device):
        adj = torch.empty((M, M), dtype=torch.float32, device=x.device)
        ws = workspace(L.gcanet_affinity_workspace_bytes(1) + 256, x.device)
        call("gcanet_affinity_matrix", ptr(x), M, C, float(sigma), ptr(adj), ptr(ws), ws.numel(), stream())
    return adj if squeeze else adj.unsqueeze(0)


def _run_ball_query(launch, n, mean_active, device):
    """The reference's retry loop (softgroup/ops/functions.py:460-472): allocate n * meanActive slots, re-run larger if the
    neighbours found do not fit; returns (idx [nActive] int32, start_len [n, 2] int32)."""
    total = torch.zeros(1, dtype=torch.int64, device=device)
    while True:
        cap = int(n) * int(mean_active)
        idx = torch.zeros(max(cap, 1), dtype=torch.int32, device=device)
        start_len = torch.zeros((n, 2), dtype=torch.int32, device=device)
        launch(idx, cap, start_len, total)
        n_active = int(total.item())
        if n_active <= cap:
            return idx[:n_active], start_len
        mean_active = n_active // n + 1


def ball_query(coords, batch_idxs, batch_offsets, adj_mat_inst, similarity_threshold_inst, adj_mat_para,
               similarity_threshold_para, radius, mean_active, with_octree=False):
    """Replaces ``ball_query`` / ``ballquery_batch_p`` (softgroup/ops/functions.py:93-104, 436-477) with the dense matrices the
    reference passes: (idx [nActive] int32, start_len [n, 2] int32).  Lists come in point order (the reference's order is
    whatever its atomicAdd produced), each in ascending neighbour index, capped at 3000 entries per point."""
    if with_octree:
        raise NotImplementedError("octree_ball_query is disabled in the reference (with_octree=False, M4:1161)")
    require_cuda(coords, "coords", torch.float32)
    require_cuda(batch_offsets, "batch_offsets", torch.int32)
    require_cuda(adj_mat_inst, "adj_mat_inst", torch.float32)
    require_cuda(adj_mat_para, "adj_mat_para", torch.float32)
    n = coords.shape[0]
    segs = batch_offsets.numel() - 1

    def launch(idx, cap, start_len, total):
        with torch.cuda.device(coords.device):
            call("gcanet_ball_query_dense", ptr(coords), ptr(batch_offsets), n, segs, ptr(adj_mat_inst),
                 float(similarity_threshold_inst), ptr(adj_mat_para), float(similarity_threshold_para), float(radius), ptr(idx),
                 cap, ptr(start_len), ptr(total), stream())

    return _run_ball_query(launch, n, mean_active, coords.device)


def affinity_ball_query(coords, batch_offsets, feat_inst, similarity_threshold_inst, feat_para, similarity_threshold_para,
                        radius, mean_active=300, sigma=1.0):
    """The fused form of the three calls at M4:1215-1233 -- two ``compute_batch_adjacency_matrix`` and the gated
    ``ball_query`` -- without the two n x n matrices: coords [n, 3], feat_inst [n, Ci], feat_para [n, Cp], batch_offsets
    [S + 1] int32 (each segment is normalised by its own largest pairwise distance, i.e. one reference call per segment).
    Returns (idx, start_len) like ``ball_query``."""
    for t, nme in ((coords, "coords"), (feat_inst, "feat_inst"), (feat_para, "feat_para")):
        require_cuda(t, nme, torch.float32)
    require_cuda(batch_offsets, "batch_offsets", torch.int32)
    n = coords.shape[0]
    segs = batch_offsets.numel() - 1
    sizes = (batch_offsets[1:] - batch_offsets[:-1])
    max_seg = int(sizes.max().item())
    L = _cabi.lib()
    ws = workspace(L.gcanet_affinity_workspace_bytes(segs), coords.device)

    def launch(idx, cap, start_len, total):
        with torch.cuda.device(coords.device):
            call("gcanet_affinity_ball_query", ptr(coords), ptr(batch_offsets), n, segs, max_seg, ptr(feat_inst), feat_inst.shape[1],
                 float(similarity_threshold_inst), ptr(feat_para), feat_para.shape[1], float(similarity_threshold_para),
                 float(sigma), float(radius), ptr(idx), cap, ptr(start_len), ptr(total), ptr(ws), ws.numel(), stream())

    return _run_ball_query(launch, n, mean_active, coords.device)
