"""ctypes front end of oracle/native_oracle.c plus the Python-level wrappers the
reference puts on top of its native ops.  TEST INFRASTRUCTURE ONLY.

Restates (paths relative to /root/reference):
  * ``knn(ref, query, k)``            models/KNN_CUDA/knn_cuda/__init__.py:41-44 (1-based -> 0-based)
  * ``KNN(k, transpose_mode)``        models/KNN_CUDA/knn_cuda/__init__.py:54-74 (python loop over the batch)
  * ``grouping_operation``            .../pointnet2_ops/pointnet2_utils.py:194-240
  * ``knn_point`` / ``group_points``  models/search_knn.py:11-14, :23-39
  * ``SoftProjection``                models/search_knn.py:44-174
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle_native.so")
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        fp = ctypes.POINTER(ctypes.c_float)
        L.oracle_knn_device.argtypes = [fp, ctypes.c_int, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        fp, ctypes.POINTER(ctypes.c_int64)]
        L.oracle_knn_device.restype = ctypes.c_int
        ip = ctypes.POINTER(ctypes.c_int32)
        L.oracle_group_points.argtypes = [ctypes.c_int] * 5 + [fp, ip, fp]
        L.oracle_group_points.restype = ctypes.c_int
        L.oracle_group_points_grad.argtypes = [ctypes.c_int] * 5 + [fp, ip, fp]
        L.oracle_group_points_grad.restype = ctypes.c_int
        L.oracle_ballquery_batch_p.argtypes = [ctypes.c_int, ctypes.c_float, fp, ip, ip, fp, ctypes.c_float, fp, ctypes.c_float,
                                               ip, ip]
        L.oracle_ballquery_batch_p.restype = ctypes.c_longlong
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def knn_device(ref: np.ndarray, query: np.ndarray, k: int):
    """ref [dim, Nr], query [dim, Nq] fp32 -> (dist [k, Nq], ind [k, Nq] 1-based)."""
    ref = np.ascontiguousarray(ref, dtype=np.float32)
    query = np.ascontiguousarray(query, dtype=np.float32)
    dim, nr = ref.shape
    nq = query.shape[1]
    dist = np.empty((k, nq), np.float32)
    ind = np.empty((k, nq), np.int64)
    rc = lib().oracle_knn_device(_fp(ref), nr, _fp(query), nq, dim, k, _fp(dist),
                                 ind.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
    if rc != 0:
        raise RuntimeError("oracle_knn_device: bad arguments")
    return dist, ind


def knn(ref: torch.Tensor, query: torch.Tensor, k: int):
    d, i = knn_device(ref.contiguous().numpy(), query.contiguous().numpy(), k)
    return torch.from_numpy(d), torch.from_numpy(i - 1)


class KNN(torch.nn.Module):
    def __init__(self, k, transpose_mode=False):
        super().__init__()
        self.k = k
        self._t = transpose_mode

    def forward(self, ref, query):
        assert ref.size(0) == query.size(0)
        flip = (lambda t: t.transpose(0, 1).contiguous()) if self._t else (lambda t: t)
        D, I = [], []
        with torch.no_grad():
            for b in range(ref.size(0)):
                d, i = knn(flip(ref[b]).float(), flip(query[b]).float(), self.k)
                D.append(flip(d))
                I.append(flip(i))
        return torch.stack(D, 0), torch.stack(I, 0)


class _Grouping(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, idx):
        f = np.ascontiguousarray(features.detach().numpy(), dtype=np.float32)
        ix = np.ascontiguousarray(idx.numpy(), dtype=np.int32)
        b, c, n = f.shape
        _, npnt, ns = ix.shape
        out = np.empty((b, c, npnt, ns), np.float32)
        rc = lib().oracle_group_points(b, c, n, npnt, ns, _fp(f),
                                       ix.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), _fp(out))
        if rc != 0:
            raise RuntimeError("oracle_group_points: index out of range")
        ctx.save_for_backward(idx)
        ctx.n = n
        return torch.from_numpy(out)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        g = np.ascontiguousarray(grad_out.numpy(), dtype=np.float32)
        ix = np.ascontiguousarray(idx.numpy(), dtype=np.int32)
        b, c, npnt, ns = g.shape
        gp = np.empty((b, c, ctx.n), np.float32)
        lib().oracle_group_points_grad(b, c, ctx.n, npnt, ns, _fp(g),
                                       ix.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), _fp(gp))
        return torch.from_numpy(gp), None


grouping_operation = _Grouping.apply


def knn_point(group_size, point_cloud, query_cloud):
    return KNN(k=group_size, transpose_mode=False)(point_cloud, query_cloud)


def group_points(group_size, point_cloud, query_cloud, point_features=None):
    _, idx = knn_point(group_size, point_cloud, query_cloud)
    idx = idx.permute(0, 2, 1).type(torch.int32).contiguous()
    gp = grouping_operation(point_cloud, idx)
    gf = None if point_features is None else grouping_operation(point_features, idx)
    return gp, gf, idx


class SoftProjection(torch.nn.Module):
    """Soft nearest-neighbour projection over knn_point + grouping_operation."""

    def __init__(self, group_size, initial_temperature=1.0, is_temperature_trainable=True, min_sigma=1e-4):
        super().__init__()
        self._group_size = group_size
        self._temperature = torch.nn.Parameter(
            torch.tensor(initial_temperature, requires_grad=is_temperature_trainable, dtype=torch.float32))
        self._min_sigma = torch.tensor(min_sigma, dtype=torch.float32)

    def sigma(self):
        return torch.max(self._temperature ** 2, self._min_sigma.to(self._temperature.device))

    def _weights(self, grouped_points, query_cloud):
        delta = grouped_points - query_cloud.unsqueeze(-1).expand_as(grouped_points)
        dist = torch.sum(delta ** 2, dim=1, keepdim=True) / self.sigma()
        return torch.softmax(-dist, dim=3)

    def _group(self, point_cloud, query_cloud, point_features=None):
        gp, gf, _ = group_points(self._group_size, point_cloud, query_cloud, point_features)
        return gp, gf

    def project(self, point_cloud, query_cloud):
        gp, _ = self._group(point_cloud, query_cloud)
        w = self._weights(gp, query_cloud).repeat(1, 3, 1, 1)
        return torch.sum(gp * w, dim=3)

    def propagate(self, point_cloud, point_features, query_cloud):
        gp, gf = self._group(point_cloud, query_cloud, point_features)
        return torch.sum(gf * self._weights(gp, query_cloud), dim=3)

    def project_and_propagate(self, point_cloud, point_features, query_cloud):
        gp, gf = self._group(point_cloud, query_cloud, point_features)
        w = self._weights(gp, query_cloud)
        return torch.sum(gp * w, dim=3), torch.sum(gf * w, dim=3)

    def forward(self, point_cloud, query_cloud, point_features=None, action="project"):
        point_cloud, query_cloud = point_cloud.contiguous(), query_cloud.contiguous()
        if action == "project":
            return self.project(point_cloud, query_cloud)
        if action == "propagate":
            return self.propagate(point_cloud, point_features, query_cloud)
        if action == "project_and_propagate":
            return self.project_and_propagate(point_cloud, point_features, query_cloud)
        raise ValueError("action should be one of the following: 'project', 'propagate', 'project_and_propagate'")


# ---------------------------------------------------------------------------------------------------------
# dense affinity + gated ball query (front end of the proposal grouping)
# ---------------------------------------------------------------------------------------------------------
def compute_batch_adjacency_matrix(batch_point_clouds, radius=0, dist_state=True, sigma=1.0):
    """Restates ``compute_batch_adjacency_matrix`` (models/dgcnn-hais-concat-direct-4.py:210-233): pairwise Euclidean
    distances (torch.cdist), zero diagonal, normalisation by the GLOBAL min / max of the tensor, exp(-d^2 / 2 sigma^2),
    zero diagonal again.  Bit-identical to the reference text (oracle/make_golden.py asserts it)."""
    d = torch.cdist(batch_point_clouds, batch_point_clouds)
    adj = d if dist_state else (d <= radius).float()
    adj = adj - torch.diag_embed(torch.diagonal(adj, dim1=-2, dim2=-1))
    lo, hi = adj.min(), adj.max()
    adj = (adj - lo) / (hi - lo)
    adj = torch.exp(-adj ** 2 / (2 * sigma ** 2))
    return adj - torch.diag_embed(torch.diagonal(adj, dim1=-2, dim2=-1))


def ball_query(coords, batch_idxs, batch_offsets, adj_mat_inst, thr_inst, adj_mat_para, thr_para, radius, mean_active=300):
    """Restates ``ball_query`` -> ``ballquery_batch_p`` (softgroup/ops/functions.py:93-104, 436-477; kernel
    bfs_cluster.cu:18-77) on CPU tensors: (idx [nActive] int32, start_len [n, 2] int32), lists in point order."""
    xyz = np.ascontiguousarray(coords.numpy(), np.float32)
    bi = np.ascontiguousarray(batch_idxs.numpy(), np.int32)
    bo = np.ascontiguousarray(batch_offsets.numpy(), np.int32)
    ai = np.ascontiguousarray(adj_mat_inst.numpy(), np.float32)
    ap = np.ascontiguousarray(adj_mat_para.numpy(), np.float32)
    n = xyz.shape[0]
    ipt = ctypes.POINTER(ctypes.c_int32)
    sl = np.zeros((n, 2), np.int32)
    L = lib()
    total = L.oracle_ballquery_batch_p(n, float(radius), _fp(xyz), bi.ctypes.data_as(ipt), bo.ctypes.data_as(ipt), _fp(ai),
                                       float(thr_inst), _fp(ap), float(thr_para), None, sl.ctypes.data_as(ipt))
    idx = np.zeros(max(int(total), 1), np.int32)
    L.oracle_ballquery_batch_p(n, float(radius), _fp(xyz), bi.ctypes.data_as(ipt), bo.ctypes.data_as(ipt), _fp(ai),
                               float(thr_inst), _fp(ap), float(thr_para), idx.ctypes.data_as(ipt), sl.ctypes.data_as(ipt))
    return torch.from_numpy(idx[:int(total)]), torch.from_numpy(sl)
