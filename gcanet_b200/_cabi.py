"""ctypes binding of libgcanet_b200.so (include/gcanet_b200.h).

The library is the product; this file only marshals torch tensors into the C-ABI's
plain pointers.  There is no fallback of any kind: if the shared library is missing
or a tensor is not on a CUDA device the call raises.
"""
from __future__ import annotations

import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libgcanet_b200.so")

c_int, c_size_t, c_void_p, c_float = ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_float


class EdgeConvDesc(ctypes.Structure):
    _fields_ = [("B", c_int), ("N", c_int), ("C", c_int), ("ldx", c_int), ("Cout", c_int), ("k", c_int),
                ("groups", c_int), ("eps", c_float), ("slope", c_float), ("storage_bf16", c_int)]


_DESC_P = ctypes.POINTER(EdgeConvDesc)


class NormalEdgeDesc(ctypes.Structure):
    _fields_ = [("B", c_int), ("N", c_int), ("ldx", c_int), ("Cout", c_int), ("k", c_int), ("groups", c_int),
                ("eps", c_float), ("slope", c_float)]


_NDESC_P = ctypes.POINTER(NormalEdgeDesc)


class GlobalFeatureDesc(ctypes.Structure):
    _fields_ = [("B", c_int), ("N", c_int), ("K", c_int), ("Cout", c_int), ("groups", c_int), ("eps", c_float)]


_GDESC_P = ctypes.POINTER(GlobalFeatureDesc)


class GroupNormDesc(ctypes.Structure):
    _fields_ = [("B", c_int), ("C", c_int), ("N", c_int), ("groups", c_int), ("eps", c_float), ("act", c_int)]


_GNDESC_P = ctypes.POINTER(GroupNormDesc)


class OffsetDesc(ctypes.Structure):
    _fields_ = [("B", c_int), ("N", c_int), ("S", c_int), ("k", c_int), ("E", c_int), ("groups", c_int),
                ("eps", c_float), ("slope", c_float)]


_ODESC_P = ctypes.POINTER(OffsetDesc)


class PrepareDesc(ctypes.Structure):
    _fields_ = [("B", c_int), ("n_raw", c_int), ("n_sub", c_int), ("max_labels", c_int), ("min_points", c_int),
                ("num_primitives", c_int)]


_PDESC_P = ctypes.POINTER(PrepareDesc)

# name -> (restype, argtypes); mirrors include/gcanet_b200.h one to one
SIGNATURES = {
    "gcanet_abi_version": (c_int, []),
    "gcanet_last_error": (ctypes.c_char_p, []),
    "gcanet_status_string": (ctypes.c_char_p, [c_int]),
    "gcanet_launch_count": (ctypes.c_ulonglong, []),
    "gcanet_check_device": (c_int, []),
    "gcanet_cn_to_nc": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "gcanet_cn_to_nc_add": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "gcanet_nc_to_cn": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "gcanet_knn_graph_columns": (c_int, [c_int, c_int]),
    "gcanet_knn_graph_workspace_bytes": (c_size_t, [c_int] * 5),
    "gcanet_knn_graph": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                 c_void_p, c_size_t, c_void_p]),
    "gcanet_knn_probe_arm": (c_int, [c_int]),
    "gcanet_knn_probe_read": (c_int, [ctypes.POINTER(c_float)]),
    "gcanet_knn_cuda_workspace_bytes": (c_size_t, [c_int] * 5),
    "gcanet_knn_cuda": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                c_void_p, c_size_t, c_void_p]),
    "gcanet_graph_feature_channels": (c_int, [c_int, c_int]),
    "gcanet_graph_feature_workspace_bytes": (c_size_t, [c_int] * 5),
    "gcanet_graph_feature": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                     c_size_t, c_void_p]),
    "gcanet_graph_feature_grad_workspace_bytes": (c_size_t, [c_int] * 5),
    "gcanet_graph_feature_grad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                          c_void_p, c_size_t, c_void_p]),
    "gcanet_group_points": (c_int, [c_int] * 5 + [c_void_p, c_void_p, c_void_p, c_void_p]),
    "gcanet_group_points_grad": (c_int, [c_int] * 5 + [c_void_p, c_void_p, c_void_p, c_void_p]),
    "gcanet_edgeconv_saved_bytes": (c_size_t, [_DESC_P]),
    "gcanet_edgeconv_workspace_bytes": (c_size_t, [_DESC_P]),
    "gcanet_edgeconv_forward": (c_int, [_DESC_P] + [c_void_p] * 9 + [c_size_t, c_void_p]),
    "gcanet_edgeconv_backward": (c_int, [_DESC_P] + [c_void_p] * 12 + [c_size_t, c_void_p]),
    "gcanet_normal_edgeconv_saved_bytes": (c_size_t, [_NDESC_P]),
    "gcanet_normal_edgeconv_workspace_bytes": (c_size_t, [_NDESC_P]),
    "gcanet_normal_edgeconv_forward": (c_int, [_NDESC_P] + [c_void_p] * 9 + [c_size_t, c_void_p]),
    "gcanet_normal_edgeconv_backward": (c_int, [_NDESC_P] + [c_void_p] * 11 + [c_size_t, c_void_p]),
    "gcanet_global_feature_saved_bytes": (c_size_t, [_GDESC_P]),
    "gcanet_global_feature_workspace_bytes": (c_size_t, [_GDESC_P]),
    "gcanet_global_feature_forward": (c_int, [_GDESC_P] + [c_void_p] * 8 + [c_size_t, c_void_p]),
    "gcanet_global_feature_backward": (c_int, [_GDESC_P] + [c_void_p] * 13 + [c_size_t, c_void_p]),
    "gcanet_group_norm_workspace_bytes": (c_size_t, [_GNDESC_P]),
    "gcanet_group_norm_forward": (c_int, [_GNDESC_P] + [c_void_p] * 6 + [c_size_t, c_void_p]),
    "gcanet_group_norm_backward": (c_int, [_GNDESC_P] + [c_void_p] * 9 + [c_size_t, c_void_p]),
    "gcanet_offset_pred_saved_bytes": (c_size_t, [_ODESC_P]),
    "gcanet_offset_pred_workspace_bytes": (c_size_t, [_ODESC_P]),
    "gcanet_offset_pred_forward": (c_int, [_ODESC_P] + [c_void_p] * 14 + [c_size_t, c_void_p]),
    "gcanet_offset_pred_backward": (c_int, [_ODESC_P] + [c_void_p] * 23 + [c_size_t, c_void_p]),
    "gcanet_prepare_samples": (c_int, [_PDESC_P] + [c_void_p] * 18),
    "gcanet_affinity_workspace_bytes": (c_size_t, [c_int]),
    "gcanet_pairwise_max_distance": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "gcanet_affinity_matrix": (c_int, [c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gcanet_ball_query_dense": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_float, c_void_p, c_float, c_float, c_void_p,
                                        ctypes.c_longlong, c_void_p, c_void_p, c_void_p]),
    "gcanet_affinity_ball_query": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_float, c_void_p, c_int, c_float,
                                           c_float, c_float, c_void_p, ctypes.c_longlong, c_void_p, c_void_p, c_void_p, c_size_t,
                                           c_void_p]),
}

_lib = None
_lock = threading.Lock()


def lib() -> ctypes.CDLL:
    """Loads the library once.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing: build it with `python -m gcanet_b200.build` "
                        "(gcanet_b200 has no CPU or eager fallback)")
                L = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(L, name)          # AttributeError if the symbol is not exported
                    fn.restype = res
                    fn.argtypes = args
                if L.gcanet_abi_version() != 2:
                    raise RuntimeError("libgcanet_b200.so ABI version mismatch")
                _lib = L
    return _lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        L = lib()
        msg = L.gcanet_last_error().decode() or L.gcanet_status_string(status).decode()
        raise RuntimeError(f"gcanet_b200 {what}: {msg} (status {status})")


def require_cuda(t: torch.Tensor, name: str, dtype=None, contiguous: bool = True) -> None:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be on CUDA: gcanet_b200 has no CPU path")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype} (got {t.dtype})")
    if contiguous and not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")


def ptr(t) -> c_void_p:
    return c_void_p(0 if t is None else t.data_ptr())


def stream() -> c_void_p:
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def workspace(nbytes: int, device) -> torch.Tensor:
    # the caching allocator hands out 512-byte aligned blocks; the C-ABI asks for 256
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def call(name: str, *args) -> None:
    check(getattr(lib(), name)(*args), name)


def launch_count() -> int:
    """Kernels launched by the library in this process (gcanet_launch_count)."""
    return int(lib().gcanet_launch_count())
