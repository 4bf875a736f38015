"""GPU suite: the compiled PyTorch C++ extension (gcanet_b200/csrc_ext/torch_ext.cpp) against the oracle and against the
ctypes front end.  Everything here is bit-exact: the extension and the ctypes path call the same C-ABI entry points, and
the KNN_CUDA / grouping kernels reproduce the reference's arithmetic (oracle/native_oracle.c)."""
import numpy as np
import pytest
import torch

import gcanet_b200 as gb
from gcanet_b200 import native_ext
from gcanet_b200.synth import abc_like_batch
from oracle import native as nat
from tests.parity import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def ext():
    return native_ext.load()


@pytest.mark.parametrize("dim,nr,nq,k", [(3, 120, 5000, 60), (3, 6, 9, 3), (5, 1001, 333, 100), (64, 300, 77, 20)])
def test_knn_matches_the_c_oracle_and_the_reference_conventions(ext, dim, nr, nq, k):
    """knn(ref, query, k) as knn.cpp:23-56 returns it: [k, Nq] distances (Euclidean, ascending) and 1-based int64 indices,
    bit-equal to the C restatement of knn.cu; the batched form equals the per-cloud calls."""
    rs = np.random.RandomState(dim * 1000 + nr)
    ref = rs.randn(2, dim, nr).astype(np.float32)
    qry = rs.randn(2, dim, nq).astype(np.float32)
    ref[:, :, 5] = ref[:, :, 3]                                       # duplicated reference points: tie rule
    R, Q = torch.from_numpy(ref).to(DEV), torch.from_numpy(qry).to(DEV)
    for b in range(2):
        dist, ind = ext.knn(R[b].contiguous(), Q[b].contiguous(), k)
        assert dist.shape == (k, nq) and ind.shape == (k, nq) and ind.dtype == torch.int64
        d, i = nat.knn_device(ref[b], qry[b], k)
        assert np.array_equal(ind.cpu().numpy(), i)                   # 1-based, like the reference
        assert np.array_equal(dist.cpu().numpy(), d)
    db, ib = torch.ops.gcanet_b200_native.knn(R, Q, k, 0)             # batched, 0-based: the Python-level knn()
    D, I = gb.KNN(k)(R, Q)
    assert torch.equal(ib, I) and torch.equal(db, D)
    with pytest.raises(RuntimeError):
        ext.knn(R[0].contiguous(), Q[0].contiguous(), nr + 1)         # k > ref_nb


def test_knn_graph_equals_the_ctypes_front_end(ext):
    x = torch.from_numpy(abc_like_batch(2, 3000, seed=11)).to(DEV)
    assert torch.equal(ext.knn_graph(x, 20, 20), gb.knn(x, 20, 20))
    assert torch.equal(ext.knn_graph(x, 10, 20), gb.knn(x, 10, 20))                                  # dilation
    feats = torch.randn(2, 64, 3000, generator=torch.Generator().manual_seed(2)).to(DEV)
    assert torch.equal(torch.ops.gcanet_b200_native.knn_graph(feats, 20, 20), gb.knn(feats, 20, 20))  # tensor-core path
    with pytest.raises(RuntimeError):
        ext.knn_graph(x, 20, 4000)                                    # k2 > N raises like topk (M4:43)


def test_group_points_forward_backward(ext):
    g = torch.Generator().manual_seed(3)
    feats = torch.randn(3, 7, 301, generator=g)
    idx = torch.randint(0, 301, (3, 57, 9), generator=g, dtype=torch.int32)
    fo = feats.clone().requires_grad_(True)
    oo = nat.grouping_operation(fo, idx)
    out = ext.group_points(feats.to(DEV), idx.to(DEV))
    assert torch.equal(out.cpu(), oo.detach())
    cot = torch.randn(oo.shape, generator=g)
    (oo * cot).sum().backward()
    gp = ext.group_points_grad(cot.to(DEV), idx.to(DEV), 301)
    assert gp.shape == (3, 7, 301) and rel_err(gp, fo.grad) < 1e-5
    # the dispatcher operator is differentiable (GroupingOperation, PN2/pointnet2_utils.py:194-240)
    fg = feats.to(DEV).requires_grad_(True)
    og = torch.ops.gcanet_b200_native.group_points(fg, idx.to(DEV))
    (og * cot.to(DEV)).sum().backward()
    assert torch.equal(og.detach().cpu(), oo.detach()) and rel_err(fg.grad, fo.grad) < 1e-5


def test_runs_on_the_current_stream_and_is_graph_capturable(ext):
    """Like the reference bindings (knn.cpp:41, group_points_gpu.cu:33) the calls are enqueued on the current stream
    without a host synchronisation -- so a call can be captured into a CUDA graph and replayed on new data."""
    g = torch.Generator().manual_seed(5)
    ref, qry = torch.randn(3, 500, generator=g).to(DEV), torch.randn(3, 2000, generator=g).to(DEV)
    static_q = qry.clone()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        ext.knn(ref, static_q, 8)
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        d_g, i_g = ext.knn(ref, static_q, 8)
    static_q.copy_(qry.flip(1))
    graph.replay()
    torch.cuda.synchronize()
    d_e, i_e = ext.knn(ref, qry.flip(1).contiguous(), 8)
    assert torch.equal(i_g, i_e) and torch.equal(d_g, d_e)
