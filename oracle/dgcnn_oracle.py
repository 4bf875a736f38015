"""CPU oracle for the DGCNN kNN-graph + EdgeConv hot path of hay-001/GCANet.

TEST INFRASTRUCTURE ONLY.  Nothing under ``gcanet_b200/`` may import this
module; it is the checker used by ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

It restates, with plain fp32 torch CPU ops, the arithmetic of the reference's
torch path (``M4`` = ``models/dgcnn-hais-concat-direct-4.py``; the same text is
duplicated in ``models/dgcnn-hais-concat-direct-2.py:30-205``):

  * ``knn``                             M4:30-47
  * ``knn_points_normals``              M4:50-90
  * ``get_graph_feature``               M4:93-124
  * ``get_graph_feature_with_normals``  M4:127-161
  * ``get_graph_feature_with_normals_g`` M4:164-205
  * ``splinenet_knn`` / ``splinenet_get_graph_feature``  models/splinenet.py:9-53
  * EdgeConv block (Conv2d 1x1 no-bias -> GroupNorm -> LeakyReLU(0.2) -> max over k)
                                        M4:469-481, M4:494-505
  * ``DGCNNEncoderGn``                  M4:455-534
  * ``conv_normal`` head                M4:584-587, M4:691-693

Parity pin: ``oracle/make_golden.py`` executes the reference's own source text
(in the build container, where ``/root/reference`` exists) on seeded inputs and
asserts that every function here returns bit-identical tensors; the resulting
vectors are committed under ``tests/golden/`` and re-checked by
``tests/test_oracle_golden.py`` on every run.  The third-party arithmetic
underneath (ATen matmul/topk/group_norm) is the container's torch 2.11 CPU
build; the reference pins torch 1.7/1.9 (requirements.txt:17, README.md:7-9).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

LEAKY_SLOPE = 0.2
GN_EPS = 1e-5


# --------------------------------------------------------------------------
# kNN (a1, a2)
# --------------------------------------------------------------------------
def _dilation_columns(k1: int, k2: int) -> np.ndarray:
    # M4:32 -- np.arange(0, k2, k2 // k1); every call site uses k1 == k2.
    return np.arange(0, k2, k2 // k1)


def neg_sqdist_matrix(xb: torch.Tensor) -> torch.Tensor:
    """One cloud, ``xb`` is [1, C, N].  Returns [1, N, N] = -|x_i - x_j|^2 in the
    reference's expansion form and operation order (M4:36-38)."""
    gram_m2 = -2 * torch.matmul(xb.transpose(2, 1), xb)
    sq = torch.sum(xb ** 2, dim=1, keepdim=True)
    return -sq - gram_m2 - sq.transpose(2, 1)


def points_normals_matrix(xb: torch.Tensor) -> torch.Tensor:
    """One cloud [1, 6, N] -> [1, N, N] of -(d_p * (1 + d_n)) (M4:61-80)."""
    pts = xb[:, 0:3]
    nrm = xb[:, 3:6]
    g = 2 * torch.matmul(pts.transpose(2, 1), pts)
    sq = torch.sum(pts ** 2, dim=1, keepdim=True)
    d_p = sq - g + sq.transpose(2, 1)
    g = 2 * torch.matmul(nrm.transpose(2, 1), nrm)
    d_n = 2 - g
    return -(d_p * (1 + d_n))


def _topk_rows(score: torch.Tensor, k1: int, k2: int) -> torch.Tensor:
    cols = _dilation_columns(k1, k2)
    return score.topk(k=k2, dim=-1)[1][:, :, cols]


def knn(x: torch.Tensor, k1: int, k2: int) -> torch.Tensor:
    """x [B, C, N] fp32 -> idx [B, N, k1] int64, nearest first, self included."""
    with torch.no_grad():
        per_cloud = [neg_sqdist_matrix(x[b:b + 1]) for b in range(x.shape[0])]
        score = torch.stack(per_cloud, 0).squeeze(1)
        return _topk_rows(score, k1, k2)


def knn_points_normals(x: torch.Tensor, k1: int, k2: int) -> torch.Tensor:
    with torch.no_grad():
        per_cloud = [points_normals_matrix(x[b:b + 1]) for b in range(x.shape[0])]
        score = torch.stack(per_cloud, 0).squeeze(1)
        return _topk_rows(score, k1, k2)


def splinenet_knn(x: torch.Tensor, k: int) -> torch.Tensor:
    """models/splinenet.py:9-22 -- same metric, single k."""
    return knn(x, k, k)


def knn_scores(x: torch.Tensor, metric: str = "l2") -> torch.Tensor:
    """[B, N, N] score matrix the reference feeds to topk (larger = nearer).
    Used by the tests to evaluate the tie tolerance on rows whose sets differ."""
    fn = neg_sqdist_matrix if metric == "l2" else points_normals_matrix
    with torch.no_grad():
        return torch.stack([fn(x[b:b + 1]) for b in range(x.shape[0])], 0).squeeze(1)


# --------------------------------------------------------------------------
# graph features (a3, a4, a5)
# --------------------------------------------------------------------------
def _gather_neighbours(x: torch.Tensor, idx: torch.Tensor, k1: int):
    """x [B, C, N], idx [B, N, k1] -> (nbr [B, N, k1, C], ctr [B, N, 1, C])."""
    B, C, N = x.shape
    flat = (idx + torch.arange(0, B, device=x.device).view(-1, 1, 1) * N).view(-1)
    pm = x.transpose(2, 1).contiguous()                  # [B, N, C]
    nbr = pm.view(B * N, -1)[flat, :].view(B, N, k1, C)
    return nbr, pm.view(B, N, 1, C)


def get_graph_feature(x, k1=20, k2=20, idx=None):
    B, N = x.size(0), x.size(2)
    x = x.view(B, -1, N)
    if idx is None:
        idx = knn(x, k1=k1, k2=k2)
    nbr, ctr = _gather_neighbours(x, idx, k1)
    ctr = ctr.repeat(1, 1, k1, 1)
    return torch.cat((nbr - ctr, ctr), dim=3).permute(0, 3, 1, 2)


def get_graph_feature_with_normals(x, k1=20, k2=20, idx=None):
    B, N = x.size(0), x.size(2)
    x = x.view(B, -1, N)
    if idx is None:
        idx = knn_points_normals(x, k1=k1, k2=k2)
    nbr, ctr = _gather_neighbours(x, idx, k1)
    ctr = ctr.repeat(1, 1, k1, 1)
    return torch.cat((nbr - ctr, ctr), dim=3).permute(0, 3, 1, 2)


def get_graph_feature_with_normals_g(x, k1=20, k2=20, idx=None):
    """[B, 6, N] -> [B, 7, N, k]: (clamp(n_i . n_j, +-0.99), n_j - n_i, n_i)  (M4:189-204)."""
    B, N = x.size(0), x.size(2)
    x = x.view(B, -1, N)
    if idx is None:
        idx = knn_points_normals(x, k1=k1, k2=k2)
    nbr, ctr = _gather_neighbours(x, idx, k1)            # [B,N,k,6], [B,N,1,6]
    n_i = ctr[..., 3:6]
    n_j = nbr[..., 3:6]
    # reference: product of [B,3,N,1] and [B,3,N,k] summed over the channel dim
    cosang = (n_i.permute(0, 3, 1, 2) * n_j.permute(0, 3, 1, 2)).sum(1).clamp(-0.99, 0.99)
    n_i = n_i.repeat(1, 1, k1, 1)
    return torch.cat((cosang.unsqueeze(-1), n_j - n_i, n_i), dim=3).permute(0, 3, 1, 2)


def splinenet_get_graph_feature(x, k=20, idx=None):
    """models/splinenet.py:25-53."""
    return get_graph_feature(x.contiguous(), k1=k, k2=k, idx=idx)


# --------------------------------------------------------------------------
# EdgeConv block (a6) and the encoder that stacks it
# --------------------------------------------------------------------------
def edgeconv_block(feat, weight, gamma, beta, groups=2, slope=LEAKY_SLOPE, eps=GN_EPS):
    """feat [B, 2C, N, k] -> [B, Cout, N]; weight [Cout, 2C] or [Cout, 2C, 1, 1]."""
    w4 = weight.view(weight.shape[0], weight.shape[1], 1, 1)
    y = F.conv2d(feat, w4)
    y = F.group_norm(y, groups, gamma, beta, eps)
    y = F.leaky_relu(y, slope)
    return y.max(dim=-1, keepdim=False)[0]


class DGCNNEncoderGn(nn.Module):
    """Same parameter names/shapes as M4:455-486 so state_dicts interchange."""

    def __init__(self, mode=0, nn_nb=80, input_channels=3):
        super().__init__()
        self.k = nn_nb
        self.mode = mode
        self.bn1 = nn.GroupNorm(2, 64)
        self.bn2 = nn.GroupNorm(2, 64)
        self.bn3 = nn.GroupNorm(2, 128)
        self.bn4 = nn.GroupNorm(4, 256)      # declared, never used (M4:466)
        self.bn5 = nn.GroupNorm(8, 1024)     # declared, never used (M4:467)
        c_in = input_channels * 2 if mode == 5 else input_channels
        act = nn.LeakyReLU(negative_slope=LEAKY_SLOPE)
        self.conv1 = nn.Sequential(nn.Conv2d(c_in, 64, kernel_size=1, bias=False), self.bn1, act)
        self.conv2 = nn.Sequential(nn.Conv2d(128, 64, kernel_size=1, bias=False), self.bn2, act)
        self.conv3 = nn.Sequential(nn.Conv2d(128, 128, kernel_size=1, bias=False), self.bn3, act)
        self.mlp1 = nn.Conv1d(256, 1024, 1)
        self.bnmlp1 = nn.GroupNorm(8, 1024)

    def edge_stack(self, x):
        """The three EdgeConv layers only: returns (x1, x2, x3)."""
        first = get_graph_feature_with_normals if self.mode == 5 else get_graph_feature
        x1 = self.conv1(first(x, k1=self.k, k2=self.k)).max(dim=-1)[0]
        x2 = self.conv2(get_graph_feature(x1, k1=self.k, k2=self.k)).max(dim=-1)[0]
        x3 = self.conv3(get_graph_feature(x2, k1=self.k, k2=self.k)).max(dim=-1)[0]
        return x1, x2, x3

    def tail(self, x1, x2, x3):
        B, N = x1.shape[0], x1.shape[2]
        cat = torch.cat((x1, x2, x3), dim=1)
        g = F.relu(self.bnmlp1(self.mlp1(cat))).max(dim=2)[0]
        g = g.view(B, 1024, 1).repeat(1, 1, N)
        return torch.cat([g, cat], 1)

    def forward(self, x):
        return self.tail(*self.edge_stack(x))


class NormalEdgeHead(nn.Module):
    """conv_normal (M4:584-587) applied as at M4:691-693."""

    def __init__(self, nn_nb=80):
        super().__init__()
        self.k = nn_nb
        self.bn_normal = nn.GroupNorm(2, 64)
        self.conv_normal = nn.Sequential(nn.Conv2d(7, 64, kernel_size=1, bias=False),
                                         self.bn_normal,
                                         nn.LeakyReLU(negative_slope=LEAKY_SLOPE))

    def forward(self, points, idx=None):
        f = get_graph_feature_with_normals_g(points, k1=self.k, k2=self.k, idx=idx)
        return self.conv_normal(f).max(dim=-1, keepdim=False)[0]


# --------------------------------------------------------------------------
# offset-prediction block: cos_dist M4:326-342, KPAM M4:351-373,
# OFFSET_PRED_MODULE M4:376-452
# --------------------------------------------------------------------------
def cos_dist(feat: torch.Tensor, keys: torch.Tensor) -> torch.Tensor:
    """feat [B, N, C], keys [B, K, C] -> [B, N, K] = -(1 - cosine similarity)  (larger = more similar)."""
    fn = feat / feat.norm(dim=-1, keepdim=True)
    kn = keys / keys.norm(dim=-1, keepdim=True)
    return -(1 - torch.einsum('bnc,bkc->bnk', fn, kn))


def offset_key_indices(num_points: int, count: int) -> torch.Tensor:
    """The reference re-seeds numpy with 1234 on every call, shuffles arange(N) and keeps the first `count`
    (M4:403-406): the same `count` points of every cloud, every step.  RandomState(1234) is the same MT19937 stream as
    the seeded global generator, without the side effect on it."""
    order = np.arange(num_points)
    np.random.RandomState(1234).shuffle(order)
    return torch.from_numpy(order[:count].copy()).long()


class KPAM(nn.Module):
    """Attention over the k neighbours from their similarity values: softmax_k(W2 relu(W1 d))."""

    def __init__(self, C):
        super().__init__()
        self.dim = C
        self.conv1 = nn.Sequential(nn.Conv1d(C, C, kernel_size=1, bias=False), nn.ReLU(),
                                   nn.Conv1d(C, C, kernel_size=1, bias=False))

    def forward(self, x, sims):
        """x [B, N, k, F], sims [B, N, k] -> x scaled per (point, neighbour)."""
        w = self.conv1(sims.permute(0, 2, 1)).permute(0, 2, 1)
        w = torch.softmax(w, dim=2).unsqueeze(-1)
        return w * x


class OffsetPredModule(nn.Module):
    """Same parameter names / shapes as OFFSET_PRED_MODULE (M4:376-452): ``bn1``, ``conv1.0.weight`` [128, 131, 1, 1],
    ``attention.conv1.{0,2}.weight`` [k, k, 1], ``mlp_offset.{weight,bias}`` [3, 256, 1] / [3]."""

    def __init__(self, nn_nb=30, sampling_ratio=120):
        super().__init__()
        self.k = nn_nb
        self.sampling_ratio = sampling_ratio
        self.bn1 = nn.GroupNorm(2, 128)
        self.conv1 = nn.Sequential(nn.Conv2d(131, 128, kernel_size=1, bias=False), self.bn1,
                                   nn.LeakyReLU(negative_slope=LEAKY_SLOPE))
        self.attention = KPAM(nn_nb)
        self.mlp_offset = nn.Conv1d(256, 3, 1)

    def forward(self, points, feature, instance_feature):
        """points [B, N, 3], feature [B, N, 128], instance_feature [B, N, E] -> offsets [B, 3, N]."""
        B, N, _ = points.shape
        sub = offset_key_indices(N, self.sampling_ratio).to(points.device)
        key_points = points[:, sub]                                  # [B, S, 3]
        key_feat = feature[:, sub]                                   # [B, S, 128]
        key_inst = instance_feature[:, sub]                          # [B, S, E]
        sims = cos_dist(instance_feature, key_inst)                  # [B, N, S]
        top_val, top_idx = torch.topk(sims, self.k, dim=2, largest=True)
        # gather through an N-fold repeat of the key tables, as the reference does (M4:425-430): its backward sums the
        # N copies in that order, which is what makes the gradients bit-identical to the reference's
        def pick(table):                                             # [B, S, F] -> [B, N, k, F]
            rep = table.unsqueeze(1).repeat(1, N, 1, 1)
            return torch.gather(rep, 2, top_idx.unsqueeze(-1).expand(-1, -1, -1, table.shape[2]))
        nb_points = pick(key_points)                                 # [B, N, k, 3]
        nb_feat = pick(key_feat)                                     # [B, N, k, 128]
        edge = torch.cat([nb_feat, nb_points - points.unsqueeze(2)], 3)          # [B, N, k, 131]
        edge = self.attention(edge, top_val)
        y = self.conv1(edge.permute(0, 3, 2, 1))                     # [B, 128, k, N]
        y = y.max(dim=-2, keepdim=False)[0]                          # [B, 128, N]
        y = torch.cat([y, feature.permute(0, 2, 1)], dim=1)          # [B, 256, N]
        return self.mlp_offset(y)
