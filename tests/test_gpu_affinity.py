"""GPU suite, part 6: dense affinity + gated ball query (front end of the proposal grouping; M4:210-233, M4:1215-1233,
softgroup/ops/src/bfs_cluster/bfs_cluster.cu:18-77).

* ``ball_query`` on the dense matrices the reference would pass: index work, BIT-EXACT against the C restatement of the
  reference kernel (lists in point order, ascending neighbour index, 3000 cap).
* ``compute_batch_adjacency_matrix``: within 2e-4 absolute of the reference formula (torch.cdist's matmul-based distances
  carry ~1e-4 relative noise for near points; ours are exact differences).
* the fused ``affinity_ball_query`` (no n x n matrix): identical neighbour sets except pairs whose affinity lies within
  1e-3 of a threshold in the reference's own noisy arithmetic; their number is reported and bounded.
"""
import numpy as np
import pytest
import torch

import gcanet_b200 as gb
from gcanet_b200.synth import abc_like_batch
from oracle import native as nat

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _scene(n, seed, clusters=7, Ci=64, Cp=22):
    g = torch.Generator().manual_seed(seed)
    xyz = torch.from_numpy(abc_like_batch(1, n, seed=seed))[0].t().contiguous()               # [n, 3]
    lab = torch.randint(0, clusters, (n,), generator=g)
    ci, cp = torch.randn(clusters, Ci, generator=g), torch.randn(clusters, Cp, generator=g)
    f_inst = ci[lab] + 0.05 * torch.randn(n, Ci, generator=g)
    f_para = cp[lab] + 0.30 * torch.randn(n, Cp, generator=g)
    return xyz, f_inst.contiguous(), f_para.contiguous()


def _lists(idx, start_len):
    idx, sl = idx.cpu().numpy(), start_len.cpu().numpy()
    return [idx[s:s + l] for s, l in sl]


@pytest.mark.parametrize("n,C", [(257, 64), (1500, 22), (3000, 64)])
def test_affinity_matrix_vs_reference_formula(n, C):
    x = torch.randn(n, C, generator=torch.Generator().manual_seed(n + C))
    want = nat.compute_batch_adjacency_matrix(x)
    got = gb.compute_batch_adjacency_matrix(x.to(DEV))
    assert got.shape == (n, n) and float(got.diagonal().abs().max()) == 0.0
    assert float((got.cpu() - want).abs().max()) < 2e-4
    assert gb.compute_batch_adjacency_matrix(x[None].to(DEV)).shape == (1, n, n)
    with pytest.raises(NotImplementedError):
        gb.compute_batch_adjacency_matrix(x.to(DEV), dist_state=False)


@pytest.mark.parametrize("n,radius,thr_i,thr_p", [(1000, 0.08, 0.989, 0.0), (3000, 0.05, 0.9, 0.5), (2000, 0.2, 0.5, -1.0),
                                                  (700, 5.0, 0.0, 0.0)])
def test_ball_query_dense_bit_exact_vs_c_restatement(n, radius, thr_i, thr_p):
    xyz, f_inst, f_para = _scene(n, seed=n)
    a_i, a_p = nat.compute_batch_adjacency_matrix(f_inst), nat.compute_batch_adjacency_matrix(f_para)
    bidx = torch.zeros(n, dtype=torch.int32)
    off = torch.tensor([0, n], dtype=torch.int32)
    idx_o, sl_o = nat.ball_query(xyz, bidx, off, a_i, thr_i, a_p, thr_p, radius)
    idx_g, sl_g = gb.ball_query(xyz.to(DEV), bidx.to(DEV), off.to(DEV), a_i.to(DEV), thr_i, a_p.to(DEV), thr_p, radius, 5)
    assert idx_g.dtype == torch.int32 and sl_g.shape == (n, 2)
    assert torch.equal(sl_g.cpu(), sl_o) and torch.equal(idx_g.cpu(), idx_o)
    if radius >= 5.0:                                           # everything is a neighbour: the 3000 cap is not reached at n = 700,
        assert int(sl_o[:, 1].max()) == n - 1                   # the zero diagonal (0 > 0 is false) drops the point itself


def test_ball_query_dense_3000_cap_and_two_segments():
    n = 7000
    g = torch.Generator().manual_seed(1)
    xyz = torch.rand(n, 3, generator=g) * 0.01                  # every pair within the radius
    ones = torch.ones(n, n)
    bidx = torch.cat([torch.zeros(4000, dtype=torch.int32), torch.ones(3000, dtype=torch.int32)])
    off = torch.tensor([0, 4000, 7000], dtype=torch.int32)
    idx_o, sl_o = nat.ball_query(xyz, bidx, off, ones, 0.5, ones, 0.5, 1.0)
    idx_g, sl_g = gb.ball_query(xyz.to(DEV), bidx.to(DEV), off.to(DEV), ones.to(DEV), 0.5, ones.to(DEV), 0.5, 1.0, 300)
    assert int(sl_o[:4000, 1].max()) == 3000 and int(sl_o[4000:, 1].min()) == 3000     # capped / whole second segment
    assert torch.equal(sl_g.cpu(), sl_o) and torch.equal(idx_g.cpu(), idx_o)


@pytest.mark.parametrize("n,radius,thr_i,thr_p", [(2000, 0.08, 0.989, 0.0), (3000, 0.06, 0.95, 0.6)])
def test_fused_affinity_ball_query_vs_reference_pipeline(n, radius, thr_i, thr_p):
    xyz, f_inst, f_para = _scene(n, seed=100 + n)
    a_i, a_p = nat.compute_batch_adjacency_matrix(f_inst), nat.compute_batch_adjacency_matrix(f_para)
    bidx = torch.zeros(n, dtype=torch.int32)
    off = torch.tensor([0, n], dtype=torch.int32)
    want = _lists(*nat.ball_query(xyz, bidx, off, a_i, thr_i, a_p, thr_p, radius))
    got = _lists(*gb.affinity_ball_query(xyz.to(DEV), off.to(DEV), f_inst.to(DEV), thr_i, f_para.to(DEV), thr_p, radius))
    borderline, pairs = 0, 0
    for i, (w, g_) in enumerate(zip(want, got)):
        pairs += len(w)
        if np.array_equal(w, g_):
            continue
        for k in np.setxor1d(w, g_):
            near = abs(float(a_i[i, k]) - thr_i) < 1e-3 or abs(float(a_p[i, k]) - thr_p) < 1e-3
            assert near, f"pair ({i}, {k}) differs away from a threshold: adj_inst {float(a_i[i, k]):.5f}, adj_para {float(a_p[i, k]):.5f}"
            borderline += 1
        assert np.all(np.diff(g_) > 0)
    print(f"n={n}: {pairs} neighbour pairs, {borderline} differ within 1e-3 of a threshold")
    assert pairs > n and borderline <= max(4, pairs // 500)


def test_fused_two_segments_and_large_cloud():
    # two clouds in one call = two reference calls (each normalised by its own largest distance)
    xa, ia, pa = _scene(1500, seed=5)
    xb, ib, pb = _scene(2500, seed=6)
    ib = ib * 3.0                                              # different feature scale: a shared normalisation would differ
    off = torch.tensor([0, 1500, 4000], dtype=torch.int32)
    both = gb.affinity_ball_query(torch.cat([xa, xb]).to(DEV), off.to(DEV), torch.cat([ia, ib]).to(DEV), 0.989,
                                  torch.cat([pa, pb]).to(DEV), 0.0, 0.08)
    la = _lists(*gb.affinity_ball_query(xa.to(DEV), off[:2].to(DEV), ia.to(DEV), 0.989, pa.to(DEV), 0.0, 0.08))
    lb = _lists(*gb.affinity_ball_query(xb.to(DEV), torch.tensor([0, 2500], dtype=torch.int32, device=DEV), ib.to(DEV), 0.989,
                                        pb.to(DEV), 0.0, 0.08))
    lists = _lists(*both)
    assert all(np.array_equal(a, b) for a, b in zip(lists[:1500], la))
    assert all(np.array_equal(a - 1500, b) for a, b in zip(lists[1500:], lb))
    # a whole 10 000-point cloud (the reference needs two 400 MB matrices for this): symmetric gates -> symmetric lists
    x, fi, fp = _scene(10000, seed=9)
    idx, sl = gb.affinity_ball_query(x.to(DEV), torch.tensor([0, 10000], dtype=torch.int32, device=DEV), fi.to(DEV), 0.989,
                                     fp.to(DEV), 0.0, 0.03)
    ls = _lists(idx, sl)
    assert int(sl[:, 1].sum()) == idx.numel() and idx.numel() > 10000
    for i in range(0, 10000, 501):
        for k in ls[i]:
            assert i in ls[int(k)]
        assert i not in ls[i]
