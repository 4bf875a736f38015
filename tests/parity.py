"""Shared parity rules.

kNN: the reference's neighbour lists come from ``topk`` over fp32 scores whose
last bits depend on the GEMM's summation order, so two correct fp32
implementations can disagree on which of two near-equidistant points is the
k-th neighbour.  The rule used everywhere:

  * per row, the index SET must equal the oracle's, except that an index the
    oracle did not pick is accepted iff its ORACLE score is within ``tau`` of the
    oracle's k-th score (then it displaced a member that is equally close);
  * the returned order must be nearest-first w.r.t. oracle scores, up to ``tau``;
  * ``tau_i = 16 * eps_fp32 * sqrt(max(C,4)/4) * (|x_i|^2 + max_j |x_j|^2)`` for
    the squared-distance metric and ``8x`` that (on the xyz part) for the
    points-x-normals metric, whose score is d_p * (1 + d_n) with 1 + d_n <= 5.

The number of rows that needed the tolerance is returned so tests can report
and bound it.
"""
from __future__ import annotations

import numpy as np
import torch

EPS32 = float(np.finfo(np.float32).eps)


def knn_tau(x: torch.Tensor, metric: str = "l2") -> torch.Tensor:
    """x [B, C, N] -> tau [B, N] (float64)."""
    xd = x.double()
    if metric == "l2":
        C = x.shape[1]
        n2 = (xd * xd).sum(1)
        scale = 16.0 * EPS32 * (max(C, 4) / 4.0) ** 0.5
    else:
        n2 = (xd[:, 0:3] ** 2).sum(1)
        scale = 128.0 * EPS32
    return scale * (n2 + n2.max(dim=1, keepdim=True)[0])


def check_knn_rows(idx_test: torch.Tensor, idx_oracle: torch.Tensor, scores: torch.Tensor,
                   tau: torch.Tensor, check_order: bool = True) -> int:
    """idx_* [B, N, k] (any int dtype), scores [B, N, N] oracle scores (larger = nearer),
    tau [B, N].  Raises AssertionError on a real mismatch; returns #rows that used tau."""
    it = idx_test.long().cpu()
    io = idx_oracle.long().cpu()
    assert it.shape == io.shape, (it.shape, io.shape)
    B, N, k = it.shape
    assert int(it.min()) >= 0 and int(it.max()) < scores.shape[-1], "index out of range"
    st = torch.gather(scores, 2, it).double()          # oracle scores of the test's picks
    so = torch.gather(scores, 2, io).double()
    kth = so.min(dim=2)[0]                             # oracle's k-th best score
    # no duplicates inside a row
    srt = it.sort(dim=2)[0]
    assert bool((srt[:, :, 1:] != srt[:, :, :-1]).all()), "duplicate neighbour in a row"
    same_set = (srt == io.sort(dim=2)[0]).all(dim=2)
    # every pick must be at least as good as the oracle's k-th, up to tau
    ok = (st >= (kth - tau).unsqueeze(-1)).all(dim=2)
    bad = ~(same_set | ok)
    if bool(bad.any()):
        b, i = [int(v[0]) for v in torch.nonzero(bad, as_tuple=True)]
        raise AssertionError(
            f"kNN row (b={b}, i={i}) differs beyond tau={float(tau[b, i]):.3e}: "
            f"test={it[b, i].tolist()} oracle={io[b, i].tolist()} "
            f"worst pick score={float(st[b, i].min()):.9g} vs oracle k-th {float(kth[b, i]):.9g}")
    if check_order:
        drop = st[:, :, 1:] - st[:, :, :-1]            # must be <= 0 (nearest first)
        assert bool((drop <= tau.unsqueeze(-1)).all()), "neighbours are not nearest-first"
    return int((~same_set).sum())


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def argmax_ambiguity(feat: torch.Tensor, weight: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, idx: torch.Tensor,
                     groups: int = 2, eps: float = 1e-5, rel: float = 2e-5):
    """Which points' input gradient is decided by a near-tie.

    EdgeConv's backward is discontinuous in two places: the max over k hands the whole gradient of (point i, channel c)
    to ONE neighbour, and LeakyReLU switches slope at 0.  Two correct fp32 implementations that round the pre-norm
    activation y differently (here: P_j + Q_i against the reference's W [x_j - x_i; x_i]) pick different neighbours when
    the two largest y of a row are closer than their rounding, and the gradient then lands on another point.  This
    helper recomputes y in fp64 and returns a bool mask [B, N] of the points whose dX can legitimately differ:

      * every neighbour j whose (sign(gamma)-oriented) y is within ``rel * max|y|`` of the row's extreme, whenever a
        row has more than one such neighbour (the gradient of that (i, c) may land on any of them);
      * every point i with a channel whose post-norm value is within ``rel * max|u|`` of the LeakyReLU kink, and the
        neighbour that attains that row's extreme.

    feat [B, 2C, N, k] (the oracle's edge tensor), weight [Cout, 2C], idx [B, N, k].  Returns (mask, n_rows_tied)."""
    B, _, N, k = feat.shape
    w = weight.reshape(weight.shape[0], -1).double()
    y = torch.einsum("oc,bcnk->bonk", w, feat.double())                        # [B, Cout, N, k]
    Cout = y.shape[1]
    sg = torch.where(gamma.double() < 0, -1.0, 1.0).view(1, Cout, 1, 1)
    z = y * sg                                                                   # the extreme that survives is max_k z
    zmax = z.max(dim=3, keepdim=True)[0]
    tol = rel * float(y.abs().max())
    near = z >= zmax - tol                                                       # [B, Cout, N, k]
    tied_rows = near.sum(dim=3) > 1                                              # [B, Cout, N]
    yg = y.view(B, groups, Cout // groups, N, k)
    mean = yg.mean(dim=(2, 3, 4), keepdim=True)
    var = yg.var(dim=(2, 3, 4), unbiased=False, keepdim=True)
    u = ((yg - mean) / torch.sqrt(var + eps)).view(B, Cout, N, k) * gamma.double().view(1, Cout, 1, 1) \
        + beta.double().view(1, Cout, 1, 1)
    u_sel = torch.gather(u, 3, z.argmax(dim=3, keepdim=True)).squeeze(3)         # [B, Cout, N]
    kink_rows = u_sel.abs() <= rel * float(u.abs().max())
    mask = torch.zeros(B, N, dtype=torch.bool)
    hit = near & (tied_rows | kink_rows).unsqueeze(3)                            # neighbours that may receive / lose it
    hit_nk = hit.any(dim=1)                                                      # [B, N, k]
    for b in range(B):
        mask[b, idx[b][hit_nk[b]]] = True
    mask |= kink_rows.any(dim=1)
    return mask, int(tied_rows.sum()) + int(kink_rows.sum())
