"""Per-point part of the GCANet embedding network: everything ``forward_train`` computes up to (and excluding) the proposal
grouping (M4:634-735) -- the part BASELINE config 4 ("full training step") can run without spconv / softgroup.ops.

Hot path and its neighbours run on the fused kernels: three EdgeConv layers + their kNN graphs, the encoder tail
(``global_feature``), the EdgeConv on normals (``normal_edgeconv``) and the offset-prediction block (``offset_pred``).
The per-point heads between them are dense 1x1 convolutions (cuBLAS through torch: consumers, SURVEY 8 out of scope) followed
by GroupNorm + ReLU on the library's streaming kernels (``group_norm_relu``: torch's GroupNorm reduces each of the B * G
2.6 MB rows with ONE CTA, 0.5 ms per call on a B200); the only other liberty taken is that the [B, 1280, N] input of ``conv1`` is never built -- its first 1024 channels are
one value per cloud, so ``W[:, :1024] x4`` is folded into a per-cloud bias.

Parameter names follow ``PrimitivesEmbeddingDGCNGn`` (M4:549-603) for the modules that exist here, so a reference checkpoint
loads with ``strict=False`` (the instance head's spconv modules have no counterpart).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as G
from .modules import OFFSET_PRED_MODULE, DGCNNEncoderGn, LEAKY_SLOPE


class PrimitivesEmbeddingPerPoint(nn.Module):
    def __init__(self, emb_size=64, num_primitives=10, mode=5, num_channels=6, nn_nb=80):
        super().__init__()
        if mode not in (0, 5):
            raise ValueError("mode 0 (xyz) or 5 (xyz + normals); mode 3 predicts normals and feeds them back (M4:680-686)")
        self.mode, self.nn_nb, self.emb_size = mode, nn_nb, emb_size
        self.encoder = DGCNNEncoderGn(mode=mode, nn_nb=nn_nb, input_channels=num_channels)
        self.offset_pred_block = OFFSET_PRED_MODULE(nn_nb=30, sampling_ratio=120)
        self.conv1 = nn.Conv1d(1024 + 256, 512, 1)
        self.bn1 = nn.GroupNorm(8, 512)
        self.conv2 = nn.Conv1d(512, 256, 1)
        self.bn2 = nn.GroupNorm(4, 256)
        self.conv3 = nn.Conv1d(262 if mode == 5 else 259, 128, 1)
        self.bn3 = nn.GroupNorm(4, 128)
        self.mlp_seg_prob1 = nn.Conv1d(832, 256, 1)
        self.mlp_seg_prob2 = nn.Conv1d(256, emb_size, 1)
        self.bn_seg_prob1 = nn.GroupNorm(4, 256)
        self.bn_normal = nn.GroupNorm(2, 64)
        self.conv_normal = nn.Sequential(nn.Conv2d(7, 64, kernel_size=1, bias=False), self.bn_normal,
                                         nn.LeakyReLU(negative_slope=LEAKY_SLOPE))
        self.mlp_prim_prob1 = nn.Conv1d(256, 256, 1)
        self.mlp_prim_prob2 = nn.Conv1d(256, num_primitives, 1)
        self.bn_prim_prob1 = nn.GroupNorm(4, 256)
        self.mlp_param_prob1 = nn.Conv1d(256, 256, 1)
        self.mlp_param_prob2 = nn.Conv1d(256, 22, 1)
        self.bn_param_prob1 = nn.GroupNorm(4, 256)

    @staticmethod
    def _gn_relu(bn, x):
        """F.relu(bn(x)) (M4:644-713) on the library's GroupNorm + ReLU kernels."""
        return G.group_norm_relu(x, bn.weight, bn.bias, bn.num_groups, bn.eps)

    @staticmethod
    def _unit(v):
        return v / (torch.norm(v, dim=-1, keepdim=True) + 1e-12)

    def forward(self, points, normals):
        """points, normals [B, N, 3] -> dict(type_per_point [B, N, P] log-probabilities, param_per_point [B, N, 22],
        pt_offsets [B, N, 3], output_feats [B, N, emb])  (the tensors the reference's per-point losses consume)."""
        B, N, _ = points.shape
        cloud = torch.cat([points, normals], dim=-1) if self.mode == 5 else points
        cloud = cloud.permute(0, 2, 1).contiguous()                                   # [B, 6 | 3, N]
        self.encoder.keep_graphs = self.mode == 5
        x4, x_feat = self.encoder.forward_global(cloud)                               # [B, 1024], [B, 256, N]
        # conv1 on cat(repeat(x4), x_features) without the repeat: the global part is a per-cloud bias
        w1 = self.conv1.weight[:, :, 0]
        bias1 = F.linear(x4, w1[:, :1024], self.conv1.bias)                           # [B, 512]
        x = F.conv1d(x_feat, w1[:, 1024:].unsqueeze(-1)) + bias1.unsqueeze(-1)
        x = self._gn_relu(self.bn1, x)
        x_all = self._gn_relu(self.bn2, self.conv2(x))
        x_type = self._gn_relu(self.bn_prim_prob1, self.mlp_prim_prob1(x_all))
        type_per_point = F.log_softmax(self.mlp_prim_prob2(x_type), dim=1).permute(0, 2, 1)
        x_para = self._gn_relu(self.bn_param_prob1, self.mlp_param_prob1(x_all))
        p = self.mlp_param_prob2(x_para).transpose(1, 2)                              # [B, N, 22]
        # sphere (4) | plane normal (3) + d | cylinder axis (3) + 4 | cone axis (3) + 4, axes normalised (M4:660-676)
        param_per_point = torch.cat([p[..., :4], self._unit(p[..., 4:7]), p[..., 7:8], self._unit(p[..., 8:11]), p[..., 11:15],
                                     self._unit(p[..., 15:18]), p[..., 18:22]], dim=2)
        # 4th EdgeConv, on normals (M4:690-693).  In mode 5 the reference recomputes the graph the encoder's first layer
        # already built (same metric, same input, M4:493 vs M4:691): reuse it
        if self.mode == 5:
            six, idx32 = cloud, self.encoder.last_graphs[0]
        else:
            six = torch.cat([cloud, normals.permute(0, 2, 1)], dim=1).contiguous()
            _, idx32 = G.knn_graph(six, self.nn_nb, self.nn_nb, G.METRIC_POINTS_NORMALS, want64=False, want32=True, ordered=False)
        gn = self.bn_normal
        normal_feature = G.normal_edgeconv(six, idx32, self.conv_normal[0].weight, gn.weight, gn.bias, groups=gn.num_groups,
                                           eps=gn.eps, slope=self.conv_normal[2].negative_slope)
        x = torch.cat([x_all, x_type, x_para, normal_feature], dim=1)                 # 256 * 3 + 64 = 832
        x = self._gn_relu(self.bn_seg_prob1, self.mlp_seg_prob1(x))
        output_feats = self.mlp_seg_prob2(x).permute(0, 2, 1)                         # [B, N, emb]
        feats_coords = torch.cat([x_all, cloud], dim=1)                               # [B, 256 + 6 | 3, N]
        feats_coords = self._gn_relu(self.bn3, self.conv3(feats_coords)).permute(0, 2, 1)    # [B, N, 128]
        pt_offsets = self.offset_pred_block(points, feats_coords.contiguous(), output_feats.contiguous()).permute(0, 2, 1)
        return {"type_per_point": type_per_point, "param_per_point": param_per_point, "pt_offsets": pt_offsets,
                "output_feats": output_feats}


def nll_loss(type_per_point, type_gt):
    """``compute_nnl_loss`` (utils/loss_utils.py:441-455): NLL over the points whose label is not -1."""
    valid = type_gt != -1
    return F.nll_loss(type_per_point[valid], type_gt[valid])


def offset_l1_loss(pt_offsets, instance_labels, offset_gt):
    """``offset_loss`` (utils/loss_utils.py:297-306): L1 over the points with an instance label, divided by their number."""
    pos = instance_labels != -1
    return F.l1_loss(pt_offsets[pos], offset_gt[pos], reduction="sum") / pos.sum().clamp_min(1)     # 0 when no point is labelled
