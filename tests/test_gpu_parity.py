"""GPU suite: the CUDA path, called through the C-ABI, against the oracle.

Tolerances (stated once, used below):
  * kNN index sets: exact, except rows whose swapped members are within ``tau`` of the
    oracle's k-th score (tests/parity.py); the count of such rows is bounded per test.
  * materialised edge features: bit-exact in fp32 for (x_j - x_i, x_i); 1e-6 abs for the
    clamped normal angle.
  * fused EdgeConv forward: |out - oracle| <= 2e-4 * max|oracle| (fp32; P_j + Q_i instead of
    W [x_j - x_i; x_i] changes rounding, GroupNorm statistics are accumulated in fp64).
  * gradients: relative 2e-3 of the largest entry (fp32 atomics sum k*N terms in arbitrary order).
"""
import json
import os

import numpy as np
import pytest
import torch

import gcanet_b200 as gb
from gcanet_b200 import functional as G
from gcanet_b200.synth import abc_like_batch
from oracle import dgcnn_oracle as orc
from oracle import native as nat
from tests.parity import check_knn_rows, knn_tau, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _t(a):
    return torch.from_numpy(np.asarray(a))


# ------------------------------------------------------------------------------ kNN
def test_knn_golden_fixture(golden_dir):
    fx = np.load(os.path.join(golden_dir, "graph_small.npz"))
    x6 = _t(fx["x6"])
    x3 = x6[:, 0:3].contiguous()
    k = int(fx["k"])
    idx = gb.knn(x3.to(DEV), k, k)
    assert idx.dtype == torch.int64 and idx.shape == (2, 257, k)
    n = check_knn_rows(idx, _t(fx["idx_l2_c3"]), orc.knn_scores(x3), knn_tau(x3))
    n += check_knn_rows(gb.knn(x3.to(DEV), 10, 20), _t(fx["idx_l2_c3_dil"]), orc.knn_scores(x3), knn_tau(x3))
    xf = _t(fx["xf"])
    n += check_knn_rows(gb.knn(xf.to(DEV), 12, 12), _t(fx["idx_l2_c64"]), orc.knn_scores(xf), knn_tau(xf))
    n += check_knn_rows(gb.knn_points_normals(x6.to(DEV), k, k), _t(fx["idx_pn"]), orc.knn_scores(x6, "pn"),
                        knn_tau(x6, "pn"))
    assert n <= 3, f"{n} rows needed the tie tolerance on the small fixture"
    # two-argument splinenet form
    assert torch.equal(gb.splinenet_knn(x3.to(DEV), k), idx)


@pytest.mark.parametrize("C,N,k,metric", [
    (3, 11, 2, "l2"), (3, 101, 10, "l2"), (3, 1001, 50, "l2"), (3, 10000, 50, "l2"), (3, 10000, 20, "l2"),
    (6, 1001, 50, "pn"), (6, 10000, 50, "pn"), (6, 10000, 80, "pn"),
    (64, 1001, 50, "l2"), (64, 10000, 50, "l2"), (64, 10000, 20, "l2"), (64, 2000, 80, "l2"),
    (128, 10000, 50, "l2"), (128, 1001, 20, "l2"), (5, 1001, 100, "l2"), (16, 777, 130, "l2"),
])
def test_knn_sweep_vs_oracle(C, N, k, metric):
    """Config 3 of BASELINE.json: C = 3/64/128, N = 10k, k = 20/50, exact-index check; plus the
    ragged sizes of the KNN_CUDA grid (11/101/1001) and the register/shared-memory list paths."""
    B = 2 if N <= 2000 else 1
    if C in (3, 6):
        x = _t(abc_like_batch(B, N, seed=100 + N, with_normals=(C == 6)))
    else:
        g = torch.Generator().manual_seed(C * 7 + N)
        x = torch.randn(B, C, N, generator=g)
    fn_o = orc.knn if metric == "l2" else orc.knn_points_normals
    fn_g = gb.knn if metric == "l2" else gb.knn_points_normals
    io = fn_o(x, k, k)
    ig = fn_g(x.to(DEV), k, k)
    n = check_knn_rows(ig, io, orc.knn_scores(x, metric), knn_tau(x, metric))
    assert n <= max(2, B * N // 500), f"{n} of {B * N} rows needed the tie tolerance"
    if metric == "l2" and C >= 64:
        # Gaussian features have well separated distances: (almost) no row may need the tolerance
        assert n <= max(3, B * N // 2000)


def test_knn_activation_features_vs_oracle():
    """Feature-space kNN on real layer-1 activations (clustered, post-LeakyReLU), not Gaussians."""
    torch.manual_seed(0)
    enc = orc.DGCNNEncoderGn(mode=0, nn_nb=20, input_channels=6)
    x = _t(abc_like_batch(1, 3000, seed=5))
    with torch.no_grad():
        x1 = enc.conv1(orc.get_graph_feature(x, 20, 20)).max(dim=-1)[0]
    io = orc.knn(x1, 50, 50)
    ig = gb.knn(x1.to(DEV), 50, 50)
    n = check_knn_rows(ig, io, orc.knn_scores(x1), knn_tau(x1))
    assert n <= 30


# ------------------------------------------------------------------------------ tensor-core kNN
@pytest.mark.parametrize("C,N,k,B", [(64, 128, 20, 1), (64, 1000, 50, 2), (64, 10000, 50, 2), (128, 10000, 50, 1),
                                     (64, 4097, 80, 2), (128, 1500, 128, 1), (64, 10000, 20, 3)])
def test_knn_tensor_core_path_vs_cuda_core_path(C, N, k, B):
    """The tcgen05 path only prunes; the survivors are re-ranked in fp32, so it must return the same
    neighbour sets as the CUDA-core scan (and the oracle) up to fp32 ties."""
    g = torch.Generator().manual_seed(C + N + k)
    x = torch.randn(B, C, N, generator=g)
    x[:, :, 7] = x[:, :, 3]                                    # exact duplicates
    xd = x.to(DEV)
    i_tc = G.knn_graph(xd, k, k, tensor_cores=True)[0]
    i_cc = G.knn_graph(xd, k, k, tensor_cores=False)[0]
    scores = orc.knn_scores(x)
    tau = knn_tau(x)
    n1 = check_knn_rows(i_tc, i_cc.cpu(), scores, tau)
    n2 = check_knn_rows(i_tc, orc.knn(x, k, k), scores, tau)
    lim = max(3, B * N // 2000)      # fp32 near-ties: a handful per 10 000 rows
    assert n1 <= lim and n2 <= lim, (n1, n2)


def test_knn_tensor_core_path_on_activations_and_ties():
    torch.manual_seed(0)
    enc = orc.DGCNNEncoderGn(mode=0, nn_nb=20, input_channels=6)
    x = _t(abc_like_batch(2, 3000, seed=5))
    with torch.no_grad():
        x1 = enc.conv1(orc.get_graph_feature(x, 20, 20)).max(dim=-1)[0]      # [2, 64, 3000], clustered
    i_tc = G.knn_graph(x1.to(DEV), 50, 50)[0]
    n = check_knn_rows(i_tc, orc.knn(x1, 50, 50), orc.knn_scores(x1), knn_tau(x1))
    assert n <= 60
    # degenerate input: every point identical -> every candidate list overflows -> CUDA-core fallback rows
    xe = torch.ones(1, 64, 600, device=DEV)
    i_e = G.knn_graph(xe, 50, 50)[0]
    assert torch.equal(i_e, G.knn_graph(xe, 50, 50, tensor_cores=False)[0])
    srt = i_e.sort(dim=2)[0]
    assert bool((srt[:, :, 1:] != srt[:, :, :-1]).all())
    # half the cloud collapsed onto one point: heavy but partial ties
    xh = torch.randn(1, 64, 2000, generator=torch.Generator().manual_seed(3))
    xh[:, :, 1000:] = xh[:, :, :1]
    i_h = G.knn_graph(xh.to(DEV), 50, 50)[0]
    check_knn_rows(i_h, orc.knn(xh, 50, 50), orc.knn_scores(xh), knn_tau(xh), check_order=True)


# ------------------------------------------------------------------------------ spatially pruned xyz kNN
@pytest.mark.parametrize("C,N,k,B,k1", [(3, 257, 20, 2, 20), (3, 10000, 50, 3, 50), (3, 10000, 80, 1, 80), (3, 5000, 20, 2, 10),
                                        (6, 10000, 50, 2, 50), (6, 3001, 80, 2, 80), (3, 100000, 50, 1, 50), (3, 4000, 150, 1, 150)])
def test_knn_xyz_pruned_path_equals_brute_force(C, N, k, B, k1):
    """Spatial sort + AABB pruning evaluates a fraction of the pairs but decides with the same fp32 distance
    arithmetic and the same (distance, index) ranking: the lists must be IDENTICAL to the brute-force scan."""
    x = _t(abc_like_batch(B, N, seed=N + k, with_normals=(C == 6)))
    x[:, :, 11] = x[:, :, 5]                                   # coincident points: index tie-break
    xd = x.to(DEV)
    metric = G.METRIC_POINTS_NORMALS if C == 6 else G.METRIC_L2
    fast = G.knn_graph(xd, k1, k, metric)[0]
    slow = G.knn_graph(xd, k1, k, metric, brute_force=True)[0]
    assert torch.equal(fast, slow)


@pytest.mark.parametrize("kind,N,k,B", [("abc", 10000, 50, 3), ("abc", 1024, 64, 2), ("abc", 6000, 100, 2), ("grid", 4096, 50, 2),
                                       ("dup", 3000, 20, 2), ("same", 2048, 50, 1), ("abc", 1100, 1, 17)])
def test_knn_xyz_both_pruned_paths_equal_brute_force(kind, N, k, B):
    """xyz clouds of 1 024 .. 32 767 points take the tensor-core scan (one K = 16 MMA per key tile: bf16 x 3 products and the
    norm, exact fp32 re-rank), ``prune=False`` keeps them on the CUDA-core kernel of knn_xyz.cu: both must return the
    brute-force lists element for element -- on surfaces, on a lattice (massive exact ties), with every point present
    four times, and with all points in one place (every row overflows into the per-row fallback)."""
    g = torch.Generator().manual_seed(N + k)
    if kind == "abc":
        x = _t(abc_like_batch(B, N, seed=N + k))
    elif kind == "grid":
        m = int(round(N ** (1 / 3))) + 1
        ax = torch.arange(m, dtype=torch.float32) / m
        x = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), 0).reshape(3, -1)[:, :N].unsqueeze(0).repeat(B, 1, 1)
    elif kind == "dup":
        x = torch.rand(B, 3, N // 4, generator=g).repeat(1, 1, 4)[:, :, torch.randperm(N, generator=g)]
    else:
        x = torch.ones(B, 3, N) * 0.3
    xd = x.contiguous().to(DEV)
    ref = G.knn_graph(xd, k, k, brute_force=True)[0]
    assert torch.equal(G.knn_graph(xd, k, k)[0], ref)
    assert torch.equal(G.knn_graph(xd, k, k, prune=False)[0], ref)
    sets = G.knn_graph(xd, k, k, want64=False, want32=True, ordered=False)[1]
    assert torch.equal(sets.sort(dim=2)[0].long(), ref.sort(dim=2)[0])


def test_knn_xyz_pruned_path_degenerate_inputs():
    # non-unit normals: no usable lower bound -> per-cloud brute-force fallback inside the same call
    x = _t(abc_like_batch(2, 2000, seed=3, with_normals=True))
    x[1, 3:6] *= 1.7
    a = G.knn_graph(x.to(DEV), 30, 30, G.METRIC_POINTS_NORMALS)[0]
    b = G.knn_graph(x.to(DEV), 30, 30, G.METRIC_POINTS_NORMALS, brute_force=True)[0]
    assert torch.equal(a, b)
    # all points identical / points on a line / a tiny cluster plus far outliers
    for xx in (torch.ones(1, 3, 700), torch.linspace(0, 1, 900).view(1, 1, 900).repeat(1, 3, 1),
               torch.cat([torch.randn(1, 3, 600) * 1e-3, torch.randn(1, 3, 40) * 50], dim=2)):
        a = G.knn_graph(xx.to(DEV), 40, 40)[0]
        b = G.knn_graph(xx.to(DEV), 40, 40, brute_force=True)[0]
        assert torch.equal(a, b)


def test_knn_errors():
    x = torch.randn(1, 3, 10)
    with pytest.raises(RuntimeError, match="no CPU path"):
        gb.knn(x, 2, 2)
    with pytest.raises(RuntimeError, match="exceeds the number of points"):
        gb.knn(x.to(DEV), 20, 20)               # topk raises in the reference (M4:43)
    with pytest.raises(RuntimeError, match="C = 6"):
        gb.knn_points_normals(x.to(DEV), 2, 2)


# ------------------------------------------------------------------------------ graph features
def test_graph_features_golden_and_grad(golden_dir):
    fx = np.load(os.path.join(golden_dir, "graph_small.npz"))
    x6 = _t(fx["x6"]).to(DEV)
    x3 = x6[:, 0:3].contiguous()
    k = int(fx["k"])
    idx = _t(fx["idx_l2_c3"]).long().to(DEV)
    f = gb.get_graph_feature(x3, k, k, idx=idx)
    assert f.shape == (2, 6, 257, k) and f.stride() == (257 * k * 6, 1, k * 6, 6)
    assert torch.equal(f.cpu(), _t(fx["gf_c3"]))
    idx_pn = _t(fx["idx_pn"]).long().to(DEV)
    assert torch.equal(gb.get_graph_feature_with_normals(x6, k, k, idx=idx_pn).cpu(), _t(fx["gf_pn"]))
    g = gb.get_graph_feature_with_normals_g(x6, k, k, idx=idx_pn)
    torch.testing.assert_close(g.cpu(), _t(fx["gf_png"]), rtol=0, atol=1e-6)
    xf = _t(fx["xf"]).to(DEV)
    ff = gb.get_graph_feature(xf, 12, 12, idx=_t(fx["idx_l2_c64"]).long().to(DEV))
    assert torch.equal(ff[:, :, ::13, :].cpu(), _t(fx["gf_c64_rows"]))
    assert torch.equal(gb.splinenet_get_graph_feature(x3, k=k, idx=idx), f)

    # gradients w.r.t. x against the oracle's autograd, both variants
    for fn_g, fn_o, xin, ii in ((gb.get_graph_feature, orc.get_graph_feature, x3, idx),
                                (gb.get_graph_feature_with_normals_g, orc.get_graph_feature_with_normals_g, x6, idx_pn)):
        xg = xin.clone().requires_grad_(True)
        xo = xin.cpu().clone().requires_grad_(True)
        og = fn_g(xg, k, k, idx=ii)
        oo = fn_o(xo, k, k, idx=ii.cpu())
        cot = torch.randn(oo.shape, generator=torch.Generator().manual_seed(1))
        (og * cot.to(DEV)).sum().backward()
        (oo * cot).sum().backward()
        assert rel_err(xg.grad, xo.grad) < 1e-5


def test_graph_feature_builds_its_own_graph():
    x = _t(abc_like_batch(2, 500, seed=9)).to(DEV)
    f = gb.get_graph_feature(x, 20, 20)
    ref = orc.get_graph_feature(x.cpu(), 20, 20, idx=gb.knn(x, 20, 20).cpu())
    assert torch.equal(f.cpu(), ref)


# ------------------------------------------------------------------------------ KNN_CUDA path
KNN_CUDA_GRID = [(400, 1000, None), (10, 100, None), (2, 10, None), (400, 1001, None), (10, 101, None), (2, 11, None),
                 (400, 30000, 50), (400, 30001, 50), (400, 10000, None), (400, 10001, None), (100, 224, None)]


@pytest.mark.parametrize("k,n,nq", KNN_CUDA_GRID)
def test_knn_cuda_vs_kdtree(k, n, nq):
    """The reference's own test (models/KNN_CUDA/tests/test_knn_cuda.py:32-87) at its own sizes:
    B=2, dim=5, transpose_mode=True, distances equal sklearn KDTree's to 3 decimals."""
    from sklearn.neighbors import KDTree
    rs = np.random.RandomState(k * 131 + n)
    ref = rs.random_sample((2, n, 5)).astype(np.float32)
    query = ref if nq is None else rs.random_sample((2, nq, 5)).astype(np.float32)
    D, I = gb.KNN(k, transpose_mode=True)(torch.from_numpy(ref).to(DEV), torch.from_numpy(query).to(DEV))
    assert D.shape == (2, query.shape[1], k) and I.dtype == torch.int64
    for b in range(2):
        dd, ii = KDTree(ref[b], leaf_size=20).query(query[b], k=k)
        np.testing.assert_almost_equal(D[b].cpu().numpy(), dd, decimal=3)
        assert (I[b].cpu().numpy() == ii).mean() > 0.99


@pytest.mark.parametrize("dim,nr,nq,k", [(3, 120, 5000, 60), (3, 6, 9, 3), (5, 1001, 333, 100), (8, 513, 100, 33),
                                         (3, 2000, 100, 400), (64, 300, 77, 20)])
def test_knn_cuda_bit_exact_vs_c_oracle(dim, nr, nq, k):
    """Same arithmetic (fma of exact differences, stable ties, sqrt): distances and indices
    must equal the C restatement of knn.cu bit for bit."""
    rs = np.random.RandomState(dim * 1000 + nr)
    ref = rs.randn(2, dim, nr).astype(np.float32)
    qry = rs.randn(2, dim, nq).astype(np.float32)
    ref[:, :, 5] = ref[:, :, 3]                          # duplicated reference points: tie rule
    D, I = gb.KNN(k)(torch.from_numpy(ref).to(DEV), torch.from_numpy(qry).to(DEV))
    for b in range(2):
        d, i = nat.knn_device(ref[b], qry[b], k)
        assert np.array_equal(I[b].cpu().numpy(), i - 1)
        assert np.array_equal(D[b].cpu().numpy(), d)
    d1, i1 = gb.knn_cuda(torch.from_numpy(ref).to(DEV), torch.from_numpy(qry).to(DEV), k, index_base=1)
    assert torch.equal(i1, I + 1) and torch.equal(d1, D)


def _reference_knn_cuda_lib():
    import ctypes
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref",
                        "libknn_cuda_ref.so")
    if not os.path.exists(path):
        return None
    lib = ctypes.CDLL(path)
    fn = getattr(lib, "_Z10knn_devicePfiS_iiiS_PlP11CUstream_st")
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    fn.restype = None
    return fn


@pytest.mark.parametrize("dim,nr,nq,k", [(3, 120, 10000, 60), (5, 1000, 1000, 100), (3, 777, 50, 400)])
def test_knn_cuda_vs_reference_kernels(dim, nr, nq, k):
    """The reference's own knn.cu, compiled unmodified into oracle/_ref by oracle/Makefile,
    run on this GPU beside ours: identical distances and indices."""
    fn = _reference_knn_cuda_lib()
    if fn is None:
        pytest.skip("oracle/_ref/libknn_cuda_ref.so not built (needs /root/reference at build time)")
    g = torch.Generator().manual_seed(dim + nr)
    ref = torch.randn(dim, nr, generator=g).to(DEV)
    qry = torch.randn(dim, nq, generator=g).to(DEV)
    dist = torch.empty(nr, nq, device=DEV)
    ind = torch.empty(k, nq, dtype=torch.int64, device=DEV)
    torch.cuda.synchronize()
    fn(ref.data_ptr(), nr, qry.data_ptr(), nq, dim, k, dist.data_ptr(), ind.data_ptr(), None)
    torch.cuda.synchronize()
    d, i = gb.knn_cuda_pair(ref, qry, k)
    assert torch.equal(i, ind - 1)
    assert torch.equal(d, dist[:k])


def test_search_knn_golden_vectors(golden_dir):
    """models/search_knn.py:180-304 -- the SoftProjection self-test's hand-written expectations."""
    with open(os.path.join(golden_dir, "search_knn_golden.json")) as f:
        gv = json.load(f)

    def bcn(a):
        return torch.tensor(a, dtype=torch.float32).t().unsqueeze(0).contiguous().to(DEV)

    for k in (1, 3):
        sp = gb.SoftProjection(k, initial_temperature=1.0).to(DEV)
        out = sp(bcn(gv["point_cloud"]), bcn(gv["query_cloud"]), bcn(gv["point_features"]), action="propagate")
        want = np.asarray(gv[f"expected_features_nn_{k}"], np.float32).T[None]
        np.testing.assert_allclose(out.detach().cpu().numpy(), want, atol=2e-3)
    sp = gb.SoftProjection(3, initial_temperature=0.1).to(DEV)       # sigma = 0.1**2 (search_knn.py:282)
    out = sp.project(bcn(gv["query_cloud"]), bcn(gv["point_cloud"]))
    np.testing.assert_allclose(out.detach().cpu().numpy(), np.asarray(gv["expected_nn_cloud"], np.float32).T[None],
                               atol=2e-3)
    # and the oracle agrees with the device path on the same calls
    spo = nat.SoftProjection(3, initial_temperature=1.0)
    o = spo.propagate(bcn(gv["point_cloud"]).cpu(), bcn(gv["point_features"]).cpu(), bcn(gv["query_cloud"]).cpu())
    spg = gb.SoftProjection(3, initial_temperature=1.0).to(DEV)
    g = spg.propagate(bcn(gv["point_cloud"]), bcn(gv["point_features"]), bcn(gv["query_cloud"]))
    torch.testing.assert_close(g.cpu(), o, rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------------------ grouping
def test_grouping_operation_forward_backward():
    g = torch.Generator().manual_seed(3)
    feats = torch.randn(3, 7, 301, generator=g)
    idx = torch.randint(0, 301, (3, 57, 9), generator=g, dtype=torch.int32)
    fo = feats.clone().requires_grad_(True)
    fg = feats.to(DEV).requires_grad_(True)
    oo = nat.grouping_operation(fo, idx)
    og = gb.grouping_operation(fg, idx.to(DEV))
    assert torch.equal(og.cpu(), oo.detach())
    cot = torch.randn(oo.shape, generator=g)
    (oo * cot).sum().backward()
    (og * cot.to(DEV)).sum().backward()
    assert rel_err(fg.grad, fo.grad) < 1e-5
    with pytest.raises(RuntimeError):
        gb.grouping_operation(fg, idx.long().to(DEV))        # CHECK_IS_INT in the reference
    with pytest.raises(RuntimeError):
        gb.grouping_operation(feats, idx)                     # "CPU not supported"
    gp, gf, ii = gb.group_points(4, feats[:, :3].contiguous().to(DEV), feats[:, :3, :50].contiguous().to(DEV),
                                 feats.to(DEV))
    assert gp.shape == (3, 3, 50, 4) and gf.shape == (3, 7, 50, 4) and ii.dtype == torch.int32
    assert torch.equal(ii[:, :, 0].cpu(), torch.arange(50, dtype=torch.int32).expand(3, 50))


@pytest.mark.parametrize("C,N", [(64, 1000), (128, 4100), (96, 260), (64, 257), (3, 1000), (6, 64), (40, 36)])
def test_to_point_major_layout(C, N):
    """[B][C][N] -> [B][N][ld] (zero-padded to a multiple of 4 channels): the 64 x 64 tile kernel (N % 4 == 0, ld >= 32,
    with partial tiles in both directions) and the 32 x 32 one for everything else; exact copy, and its autograd."""
    x = torch.randn(3, C, N, generator=torch.Generator().manual_seed(C * N)).to(DEV).requires_grad_(True)
    ld = (C + 3) // 4 * 4
    y = G._ToPointMajor.apply(x, ld)
    assert y.shape == (3, N, ld)
    assert torch.equal(y[:, :, :C], x.detach().transpose(1, 2))
    assert not y[:, :, C:].any()
    cot = torch.randn_like(y)
    (y * cot).sum().backward()
    assert torch.equal(x.grad, cot[:, :, :C].transpose(1, 2))
    # transpose + accumulate in one pass (the two incoming gradients of x1 / x2): bit-identical to the two-step sum
    add = torch.randn(3, N, ld, generator=torch.Generator().manual_seed(7)).to(DEV)
    z = G.to_point_major(x.detach(), ld, add=add)
    assert torch.equal(z, y.detach() + add)


# ------------------------------------------------------------------------------ fused EdgeConv
@pytest.mark.parametrize("C,Cout,N,k,groups", [(3, 64, 257, 20, 2), (6, 64, 300, 16, 2), (64, 64, 200, 50, 2),
                                               (64, 128, 190, 33, 2), (64, 32, 64, 8, 4), (16, 256, 70, 5, 8),
                                               (64, 128, 700, 50, 2), (64, 256, 600, 20, 4)])   # X~ scatter path
def test_edgeconv_forward_backward_vs_oracle(C, Cout, N, k, groups):
    g = torch.Generator().manual_seed(C + Cout + N)
    B = 2
    x = torch.randn(B, C, N, generator=g)
    W = torch.randn(Cout, 2 * C, generator=g) / (2 * C) ** 0.5
    gamma = torch.randn(Cout, generator=g) * 0.7 + 0.2        # both signs: exercises the max/min switch
    beta = torch.randn(Cout, generator=g) * 0.3
    idx = orc.knn(x, k, k)
    cot = torch.randn(B, Cout, N, generator=g)

    xo, Wo, go, bo = (t.clone().requires_grad_(True) for t in (x, W, gamma, beta))
    out_o = orc.edgeconv_block(orc.get_graph_feature(xo, k, k, idx=idx), Wo, go, bo, groups=groups)
    (out_o * cot).sum().backward()

    xg, Wg, gg, bg = (t.to(DEV).requires_grad_(True) for t in (x, W, gamma, beta))
    x_nc = G._ToPointMajor.apply(xg, (C + 3) // 4 * 4)
    out_nc, out_cn = gb.edgeconv(x_nc, idx.int().to(DEV), Wg, gg, bg, C, groups=groups)
    assert torch.equal(out_cn, out_nc.transpose(1, 2))
    scale = float(out_o.detach().abs().max())
    assert float((out_cn.cpu() - out_o).abs().max()) <= 2e-4 * scale
    (out_cn * cot.to(DEV)).sum().backward()
    for name, a, b in (("dx", xg.grad, xo.grad), ("dW", Wg.grad, Wo.grad), ("dgamma", gg.grad, go.grad),
                       ("dbeta", bg.grad, bo.grad)):
        assert rel_err(a, b) < 2e-3, f"{name}: rel err {rel_err(a, b):.3e}"


@pytest.mark.parametrize("Cout,N,k", [(128, 4501, 8), (64, 4200, 6)])
def test_edgeconv_backward_large_m_vs_oracle(Cout, N, k):
    """M = B * N >= 8192 rows: the weight gradient runs on the tensor cores (gemm_tn_tc_kernel) and every projection GEMM
    takes its tensor-core path.  At ~10^6 (point, channel) outputs a handful sit within fp32 rounding of a LeakyReLU kink
    or of an arg-max tie, where the gradient is discontinuous (the oracle and any fp32 implementation may land on
    different sides): dx is therefore compared per point with a bounded fraction of outliers, the summed gradients with a
    tolerance that absorbs a few such flips (one flip moves a row of dW by up to ~1e-2 of the largest entry)."""
    C, B, groups = 64, 2, 2
    g = torch.Generator().manual_seed(Cout + N)
    x = torch.randn(B, C, N, generator=g)
    W = torch.randn(Cout, 2 * C, generator=g) / (2 * C) ** 0.5
    gamma = torch.randn(Cout, generator=g) * 0.7 + 0.2
    beta = torch.randn(Cout, generator=g) * 0.3
    idx = orc.knn(x, k, k)
    cot = torch.randn(B, Cout, N, generator=g)
    xo, Wo, go, bo = (t.clone().requires_grad_(True) for t in (x, W, gamma, beta))
    out_o = orc.edgeconv_block(orc.get_graph_feature(xo, k, k, idx=idx), Wo, go, bo, groups=groups)
    (out_o * cot).sum().backward()
    xg, Wg, gg, bg = (t.to(DEV).requires_grad_(True) for t in (x, W, gamma, beta))
    out_nc, out_cn = gb.edgeconv(G._ToPointMajor.apply(xg, C), idx.int().to(DEV), Wg, gg, bg, C, groups=groups)
    assert float((out_cn.cpu() - out_o).abs().max()) <= 2e-4 * float(out_o.detach().abs().max())
    (out_cn * cot.to(DEV)).sum().backward()
    per_point = (xg.grad.cpu() - xo.grad).abs().amax(dim=1)                   # [B, N]
    outliers = float((per_point > 2e-3 * float(xo.grad.abs().max())).float().mean())
    assert outliers < 3e-3, f"dx: {outliers:.2e} of the points differ"
    for name, a, b in (("dW", Wg.grad, Wo.grad), ("dgamma", gg.grad, go.grad), ("dbeta", bg.grad, bo.grad)):
        assert rel_err(a, b) < 3e-2, f"{name}: rel err {rel_err(a, b):.3e}"
    # a flip touches one row of dW (one output channel); the rest agrees to fp32-GEMM accuracy
    dw_err = (Wg.grad.cpu() - Wo.grad).abs().flatten() / float(Wo.grad.abs().max())
    assert float(dw_err.median()) < 2e-5 and float(dw_err.quantile(0.9)) < 2e-4, (float(dw_err.median()), float(dw_err.quantile(0.9)))


@pytest.mark.parametrize("mode", [0, 5])
def test_encoder_edge_stack_golden(golden_dir, mode):
    """DGCNNEncoderGn with the fixture's weights: x1|x2|x3 and the hot-path parameter gradients
    against values produced by the reference's own source (oracle/make_golden.py)."""
    fx = np.load(os.path.join(golden_dir, "encoder_small.npz"))
    k = int(fx["k"])
    x6 = _t(fx["x6"])
    x = (x6 if mode == 5 else x6[:, 0:3]).contiguous().to(DEV)
    enc = gb.DGCNNEncoderGn(mode=mode, nn_nb=k, input_channels=6)
    sd = enc.state_dict()
    for name in list(sd):
        key = f"m{mode}.param.{name}"
        if key in fx.files:
            sd[name] = _t(fx[key])
    enc.load_state_dict(sd)
    enc.to(DEV)
    x1, x2, x3 = enc.edge_stack(x)
    out = torch.cat((x1, x2, x3), 1)
    want = _t(fx[f"m{mode}.x123"])
    # layers 2 and 3 build their graphs on computed activations: a near-tie can flip a neighbour,
    # which changes that point's max slightly; bound the fraction of such points instead of all.
    diff = (out.cpu() - want).abs()
    tol = 5e-4 * float(want.abs().max())
    assert float((diff > tol).float().mean()) < 2e-3, float(diff.max())
    assert float((out.cpu()[:, :64] - want[:, :64]).abs().max()) <= tol     # layer 1: same graph
    (out * _t(fx[f"m{mode}.cot"]).to(DEV)).sum().backward()
    params = dict(enc.named_parameters())
    for key in fx.files:
        if key.startswith(f"m{mode}.grad."):
            name = key[len(f"m{mode}.grad."):]
            assert rel_err(params[name].grad, _t(fx[key])) < 5e-3, name
    assert params["bn4.weight"].grad is None and params["bn5.weight"].grad is None
    full = enc(x)
    assert full.shape == (2, 1280, 192)
    assert torch.equal(full[:, 1024:], out.detach())


def test_encoder_matches_oracle_full_size_single_cloud():
    """Config 1 shape: one 10 000-point cloud, k = 50, mode 0, forward only."""
    torch.manual_seed(0)
    ref = orc.DGCNNEncoderGn(mode=0, nn_nb=50, input_channels=6)
    x = _t(abc_like_batch(1, 10000, seed=1234))
    with torch.no_grad():
        x1o, x2o, x3o = ref.edge_stack(x)
    enc = gb.DGCNNEncoderGn(mode=0, nn_nb=50, input_channels=6)
    enc.load_state_dict(ref.state_dict())
    enc.to(DEV)
    with torch.no_grad():
        x1, x2, x3 = enc.edge_stack(x.to(DEV))
    # a neighbour flipped by an fp32 tie (the oracle itself differs from fp64 on ~2 of 10 000 rows,
    # BASELINE.md section 2) moves that point's max: bound the fraction of affected outputs
    d1 = (x1.cpu() - x1o).abs()
    assert float((d1 > 2e-4 * float(x1o.abs().max())).float().mean()) < 2e-4
    for a, b in ((x2, x2o), (x3, x3o)):
        d = (a.cpu() - b).abs()
        assert float((d > 5e-4 * float(b.abs().max())).float().mean()) < 2e-3


# ------------------------------------------------------------------------------ full-size properties
def test_full_size_properties_b16():
    """BASELINE config 2 sizes (B=16, N=10k, k=50): properties that need no oracle."""
    B, N, k = 16, 10000, 50
    x = _t(abc_like_batch(B, N, seed=1234)).to(DEV)
    idx = gb.knn(x, k, k)
    assert idx.shape == (B, N, k) and int(idx.min()) >= 0 and int(idx.max()) < N
    # self is the nearest neighbour (distance 0) except for coincident points
    self_first = (idx[:, :, 0] == torch.arange(N, device=DEV)).float().mean()
    assert float(self_first) > 0.999
    # rows are sorted nearest-first under exact fp64 distances of the returned neighbours
    xp = x.transpose(1, 2).double()
    rows = torch.arange(0, N, 97, device=DEV)
    sel = idx[:, rows]                                              # [B, R, k]
    pts = torch.gather(xp, 1, sel.reshape(B, -1, 1).expand(B, rows.numel() * k, 3)).view(B, rows.numel(), k, 3)
    d = ((pts - xp[:, rows].unsqueeze(2)) ** 2).sum(-1)
    assert bool((d[:, :, 1:] - d[:, :, :-1] >= -1e-6).all())
    # the k-th neighbour distance upper-bounds nothing closer outside the set: check on sampled rows
    dall = ((xp[:, rows].unsqueeze(2) - xp.unsqueeze(1)) ** 2).sum(-1)      # [B, R, N]
    kth = d[:, :, -1:]
    inside = torch.zeros_like(dall, dtype=torch.bool).scatter_(2, sel, True)
    assert bool((dall[~inside].view(B, rows.numel(), N - k) >= kth - 1e-6).all())
    # idempotence / determinism
    assert torch.equal(gb.knn(x, k, k), idx)


def test_large_cloud_stress_100k():
    """Config 5 (reduced to one cloud to keep the suite short): 100 000 points, k = 50; the
    reference itself cannot run this (40 GB distance matrix per cloud, M4:36-41).  Checked
    against a chunked fp32 oracle on sampled query rows."""
    N, k = 100000, 50
    x = _t(abc_like_batch(1, N, seed=77))
    idx = gb.knn(x.to(DEV), k, k).cpu()
    rows = torch.arange(0, N, 1999)
    xb = x[0]                                                      # [3, N]
    sq = torch.sum(xb ** 2, dim=0, keepdim=True)
    score = -sq - (-2 * torch.matmul(xb[:, rows].t(), xb)) - sq[:, rows].t()      # [R, N], reference form
    io = score.topk(k, dim=-1)[1]
    tau = knn_tau(x)[0, rows]
    st = torch.gather(score, 1, idx[0, rows]).double()
    kth = torch.gather(score, 1, io).double().min(dim=1)[0]
    assert bool((st >= (kth - tau).unsqueeze(-1)).all())
    # rows whose set differs from the chunked oracle's: every one passed the tau check above; report and bound the count
    n_diff = int((idx[0, rows].sort(dim=1)[0] != io.sort(dim=1)[0]).any(dim=1).sum())
    print(f"100k xyz: {n_diff} of {rows.numel()} sampled rows needed the tie tolerance")
    assert n_diff <= max(2, rows.numel() // 25)


def test_knn_xyz_pruned_path_cloud_edges():
    """First / last Morton tiles see fewer neighbours in phase 1 (threshold may still be +inf when the
    pruned phase starts): their tiles must not be scanned twice."""
    for N, k in [(1001, 50), (2000, 41), (300, 100), (257, 150)]:
        x = _t(abc_like_batch(2, N, seed=100 + N)).to(DEV)
        a = G.knn_graph(x, k, k)[0]
        assert torch.equal(a, G.knn_graph(x, k, k, brute_force=True)[0])


@pytest.mark.parametrize("C,N,k", [(64, 10000, 50), (128, 3000, 20), (64, 700, 80), (3, 5000, 50)])
def test_knn_unordered_mode_returns_the_same_sets(C, N, k):
    """ordered=False (what the fused encoder asks for) may skip the final ordering but must return
    exactly the same neighbour set per point."""
    g = torch.Generator().manual_seed(C + N)
    x = (torch.randn(2, C, N, generator=g) if C > 3 else _t(abc_like_batch(2, N, seed=1))).to(DEV)
    a = G.knn_graph(x, k, k, want64=False, want32=True, ordered=True)[1]
    b = G.knn_graph(x, k, k, want64=False, want32=True, ordered=False)[1]
    assert torch.equal(a.sort(dim=2)[0], b.sort(dim=2)[0])
    # real activations (clustered features)
    if C == 64:
        torch.manual_seed(0)
        enc = orc.DGCNNEncoderGn(mode=0, nn_nb=20, input_channels=6)
        xa = _t(abc_like_batch(2, 3000, seed=5))
        with torch.no_grad():
            x1 = enc.conv1(orc.get_graph_feature(xa, 20, 20)).max(dim=-1)[0].to(DEV)
        a = G.knn_graph(x1, k, k, want64=False, want32=True, ordered=True)[1]
        b = G.knn_graph(x1, k, k, want64=False, want32=True, ordered=False)[1]
        assert torch.equal(a.sort(dim=2)[0], b.sort(dim=2)[0])


def test_knn_tensor_core_path_on_spatially_sorted_features():
    """Keys arriving in a spatially coherent order (sorted clouds) are the adversarial case for a
    streaming threshold; the strided tile order must keep the result exact."""
    g = torch.Generator().manual_seed(11)
    x = torch.randn(1, 64, 5000, generator=g)
    order = torch.argsort(x[0, 0])                    # sort the cloud along one feature axis
    xs = x[:, :, order].contiguous()
    i_tc = G.knn_graph(xs.to(DEV), 50, 50)[0]
    n = check_knn_rows(i_tc, orc.knn(xs, 50, 50), orc.knn_scores(xs), knn_tau(xs))
    assert n <= 3


def _layer_activations(B, N, seed, k=20):
    """x1, x2 of the oracle encoder on synthetic clouds: the clustered, low-intrinsic-dimension features
    the bounding-box pruning of the tensor-core path is built for."""
    torch.manual_seed(0)
    enc = orc.DGCNNEncoderGn(mode=0, nn_nb=k, input_channels=6)
    x = _t(abc_like_batch(B, N, seed=seed))
    with torch.no_grad():
        x1 = enc.conv1(orc.get_graph_feature(x, k, k)).max(dim=-1)[0]
        x2 = enc.conv2(orc.get_graph_feature(x1, k, k)).max(dim=-1)[0]
    return x1, x2


@pytest.mark.parametrize("N,k,B", [(10000, 50, 2), (4097, 20, 2), (1025, 64, 1), (3000, 50, 3), (6000, 80, 2), (2500, 128, 1)])
def test_knn_pruned_tensor_core_path_equals_full_scan(N, k, B):
    """Sorting the cloud along its principal directions and skipping key tiles by bounding box must not
    change a single index: same lists as the full tensor-core scan, ordered and unordered, on real
    activations (where most tiles are skipped) and on Gaussians (where none are)."""
    x1, x2 = _layer_activations(B, N, seed=21)
    g = torch.Generator().manual_seed(N + k)
    for x in (x1, x2, torch.randn(B, 64, N, generator=g)):
        xd = x.to(DEV)
        a = G.knn_graph(xd, k, k, prune=True)[0]
        b = G.knn_graph(xd, k, k, prune=False)[0]
        assert torch.equal(a, b)
        c = G.knn_graph(xd, k, k, want64=False, want32=True, ordered=False, prune=True)[1]
        assert torch.equal(c.sort(dim=2)[0].long(), a.sort(dim=2)[0])
    n = check_knn_rows(a, orc.knn(x, k, k), orc.knn_scores(x), knn_tau(x))
    assert n <= max(3, B * N // 2000)


def test_knn_pruned_path_c128_and_degenerate_clouds():
    x1, x2 = _layer_activations(1, 3000, seed=8)
    x = torch.cat([x1, x2], dim=1)                     # [1, 128, 3000]
    xd = x.to(DEV)
    assert torch.equal(G.knn_graph(xd, 50, 50)[0], G.knn_graph(xd, 50, 50, prune=False)[0])
    # rank-deficient clouds: constant, one-dimensional, and a cloud with a NaN-free huge outlier
    z = torch.ones(1, 64, 1500, device=DEV)
    assert torch.equal(G.knn_graph(z, 20, 20)[0], G.knn_graph(z, 20, 20, prune=False)[0])
    line = torch.zeros(1, 64, 2000)
    line[0, 5] = torch.linspace(-1, 1, 2000)
    assert torch.equal(G.knn_graph(line.to(DEV), 20, 20)[0], G.knn_graph(line.to(DEV), 20, 20, prune=False)[0])
    out = x1.clone()
    out[0, :, 17] += 1000.0
    od = out.to(DEV)
    assert torch.equal(G.knn_graph(od, 50, 50)[0], G.knn_graph(od, 50, 50, prune=False)[0])


@pytest.mark.parametrize("B,N,k", [(1, 1024, 1), (33, 1087, 7), (5, 2048, 64), (2, 6000, 50)])
def test_knn_pruned_path_shapes_and_ties(B, N, k):
    """Cloud counts that change the sort-key layout, sizes around the tile boundaries, k = 1 and k = 64 (every
    column slot needed), and clouds where a third of the points coincide (candidate lists overflow, rows go to
    the CUDA-core fallback list): always the same lists as the full scan."""
    x1, _ = _layer_activations(min(B, 3), N, seed=40 + B)
    x = x1.repeat((B + x1.shape[0] - 1) // x1.shape[0], 1, 1)[:B].clone()
    x += 0.01 * torch.randn(x.shape, generator=torch.Generator().manual_seed(B))      # clouds differ
    xd = x.to(DEV)
    assert torch.equal(G.knn_graph(xd, k, k)[0], G.knn_graph(xd, k, k, prune=False)[0])
    xt = x.clone()
    xt[:, :, : N // 3] = xt[:, :, :1]                      # massive exact ties
    xd = xt.to(DEV)
    a = G.knn_graph(xd, k, k)[0]
    b = G.knn_graph(xd, k, k, brute_force=True)[0]
    # tied points are interchangeable: compare the distance profile, and exact equality away from the ties
    sc = orc.knn_scores(xt[:1])
    check_knn_rows(a[:1], b[:1].cpu(), sc, knn_tau(xt[:1]))
    srt = a.sort(dim=2)[0]
    assert bool((srt[:, :, 1:] != srt[:, :, :-1]).all())


def test_knn_and_edgeconv_are_graph_capturable():
    """The C-ABI never allocates or synchronises (include/gcanet_b200.h), so a kNN + EdgeConv forward can be
    captured in a CUDA graph and replayed on new inputs of the same shape."""
    x1, _ = _layer_activations(2, 2000, seed=3)
    a, b = x1.to(DEV), x1.flip(2).contiguous().to(DEV)
    w = torch.randn(64, 128, generator=torch.Generator().manual_seed(1)).to(DEV) * 0.1
    gm, bt = torch.ones(64, device=DEV), torch.zeros(64, device=DEV)

    def run(x):
        with torch.no_grad():
            idx = G.knn_graph(x, 20, 20, want64=False, want32=True, ordered=False)[1]
            out_nc, out_cn = gb.edgeconv(G.to_point_major(x), idx, w, gm, bt, 64, groups=2)
        return idx, out_cn

    static_x = a.clone()
    for _ in range(2):
        run(static_x)                                   # warm-up: workspaces, function attributes
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        idx_g, out_g = run(static_x)
    for src in (a, b):
        static_x.copy_(src)
        g.replay()
        torch.cuda.synchronize()
        idx_e, out_e = run(src)
        assert torch.equal(idx_g.sort(dim=2)[0], idx_e.sort(dim=2)[0])
        assert torch.allclose(out_g, out_e, rtol=1e-5, atol=1e-6)


def test_scan_probe_brackets_only_tensor_core_scans():
    """gcanet_knn_probe_arm / _read (bench.py's roofline leg): a feature-space call reports a positive scan time that is
    no longer than the whole call, the neighbour lists are the same with and without the probe, a call on a path without
    a tensor-core scan (brute force) leaves the probe unfired, and a read without an armed call is an error."""
    from gcanet_b200 import _cabi
    import ctypes
    L = _cabi.lib()
    ms = ctypes.c_float(-1.0)
    L.gcanet_knn_probe_arm(0)
    assert L.gcanet_knn_probe_read(ctypes.byref(ms)) == -1
    x1, _ = _layer_activations(2, 4000, seed=5)
    x = x1.to(DEV)
    ref = G.knn_graph(x, 20, 20, want64=False, want32=True)[1]
    G.enable_kernel_timing(True)
    G.enable_scan_probe(True)
    try:
        got = G.knn_graph(x, 20, 20, want64=False, want32=True)[1]
        G.knn_graph(x, 20, 20, want64=False, want32=True, brute_force=True)
        torch.cuda.synchronize()
        calls, scans = G.kernel_timings_ms(), G.scan_kernel_timings_ms()
    finally:
        G.enable_scan_probe(False)
        G.enable_kernel_timing(False)
    assert torch.equal(got, ref)
    tag = "knn_graph[C=64,metric=0]"
    assert list(scans) == [tag] and len(scans[tag]) == 1          # the brute-force call bracketed nothing
    assert 0.0 < scans[tag][0] <= calls[tag][0]


def test_normal_edge_head_golden(golden_dir):
    """conv_normal head (M4:584-587, 691-693) against the fixture made from the reference's
    get_graph_feature_with_normals_g; forward 1e-4, weight gradients 2e-3 relative."""
    fx = np.load(os.path.join(golden_dir, "normal_head_small.npz"))
    torch.backends.cudnn.allow_tf32 = False          # the 7 -> 64 conv is torch's (cuDNN defaults to TF32)
    head = gb.NormalEdgeHead(nn_nb=int(fx["k"]))
    with torch.no_grad():
        for name, p in head.named_parameters():
            p.copy_(_t(fx[f"param.{name}"]))
    head.to(DEV)
    x6 = _t(fx["x6"]).to(DEV)
    out = head(x6)
    want = _t(fx["out"])
    d = (out.cpu() - want).abs()
    assert float((d > 1e-4 * float(want.abs().max())).float().mean()) < 2e-3      # a tie-flipped neighbour moves a max
    (out * _t(fx["cot"]).to(DEV)).sum().backward()
    for name, p in head.named_parameters():
        assert rel_err(p.grad, _t(fx[f"grad.{name}"])) < 5e-3, name
    # reusing the encoder's layer-1 graph (mode 5) gives the same result
    idx = gb.knn_points_normals(x6, int(fx["k"]), int(fx["k"]))
    assert float((head(x6, idx=idx) - out).abs().max()) < 1e-5
    # the differentiable-in-points fallback (materialised feature + torch conv) agrees with the fused kernel
    xg = x6.clone().requires_grad_(True)
    out2 = head(xg)
    d2 = (out2 - out).abs()
    assert float((d2 > 1e-4 * float(want.abs().max())).float().mean()) < 2e-3
    out2.sum().backward()
    assert xg.grad is not None and bool(torch.isfinite(xg.grad).all())


def test_large_cloud_feature_space_100k():
    """Config 5, feature-space half: one 100 000-point cloud with C = 64 through the tensor-core path
    (the reference needs a 40 GB distance matrix per cloud, M4:36-41); sampled rows against a chunked
    fp32 oracle in the reference's expansion arithmetic."""
    N, k, C = 100000, 50, 64
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, C, N, generator=g)
    idx = gb.knn(x.to(DEV), k, k).cpu()
    rows = torch.arange(0, N, 2503)
    xb = x[0]
    sq = torch.sum(xb ** 2, dim=0, keepdim=True)
    score = -sq - (-2 * torch.matmul(xb[:, rows].t(), xb)) - sq[:, rows].t()        # [R, N]
    io = score.topk(k, dim=-1)[1]
    tau = knn_tau(x)[0, rows]
    st = torch.gather(score, 1, idx[0, rows]).double()
    kth = torch.gather(score, 1, io).double().min(dim=1)[0]
    assert bool((st >= (kth - tau).unsqueeze(-1)).all())
    n_diff = int((idx[0, rows].sort(dim=1)[0] != io.sort(dim=1)[0]).any(dim=1).sum())
    print(f"100k C=64: {n_diff} of {rows.numel()} sampled rows needed the tie tolerance")
    assert n_diff <= 1                                    # Gaussian features: distances are well separated
    srt = idx.sort(dim=2)[0]
    assert bool((srt[:, :, 1:] != srt[:, :, :-1]).all())
