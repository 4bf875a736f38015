"""CPU suite, part 3: the compiled PyTorch C++ extension over the C-ABI (gcanet_b200/csrc_ext/torch_ext.cpp).

No compute without a GPU: the module imports, exposes the reference's pybind names (knn.cpp:59-61, bindings.cpp:17-18),
registers its dispatcher operators with the documented schemas, infers shapes on fake tensors, and refuses CPU tensors
the way the reference does (CHECK_INPUT at knn.cpp:5-8,29-30; "CPU not supported" at group_points.cpp:32,58).
"""
import pytest
import torch

from gcanet_b200 import native_ext


@pytest.fixture(scope="module")
def ext():
    return native_ext.load()


def test_module_exposes_the_reference_names(ext):
    for name in ("knn", "group_points", "group_points_grad", "knn_graph"):
        assert callable(getattr(ext, name))
    assert ext.abi_version() == 2
    assert native_ext.load() is ext                                  # loaded once


def test_cpu_tensors_and_wrong_dtypes_raise_like_the_reference(ext):
    with pytest.raises(RuntimeError, match="must be on CUDA"):
        ext.knn(torch.zeros(3, 5), torch.zeros(3, 4), 2)
    with pytest.raises(RuntimeError, match="must be at::kFloat"):
        ext.knn(torch.zeros(3, 5, dtype=torch.float64), torch.zeros(3, 4), 2)
    with pytest.raises(RuntimeError, match="must be contiguous"):
        ext.knn(torch.zeros(5, 3).t(), torch.zeros(3, 4), 2)
    with pytest.raises(RuntimeError, match="CPU not supported"):
        ext.group_points(torch.zeros(1, 3, 5), torch.zeros(1, 2, 2, dtype=torch.int32))
    with pytest.raises(RuntimeError, match="CPU not supported"):
        ext.group_points_grad(torch.zeros(1, 3, 2, 2), torch.zeros(1, 2, 2, dtype=torch.int32), 5)
    with pytest.raises(RuntimeError, match="must be an int tensor"):
        ext.group_points(torch.zeros(1, 3, 5), torch.zeros(1, 2, 2, dtype=torch.int64))
    with pytest.raises(RuntimeError, match="must be on CUDA"):
        ext.knn_graph(torch.zeros(1, 3, 5), 2, 2)


def test_dispatcher_ops_have_no_cpu_kernel_and_infer_shapes(ext):
    ns = torch.ops.gcanet_b200_native
    assert str(ns.knn.default._schema) == \
        "gcanet_b200_native::knn(Tensor ref, Tensor query, int k, int index_base=1) -> (Tensor, Tensor)"
    assert str(ns.knn_graph.default._schema) == "gcanet_b200_native::knn_graph(Tensor x, int k1, int k2, int metric=0) -> Tensor"
    with pytest.raises(NotImplementedError):                         # CUDA key only: no fallback
        ns.knn(torch.zeros(3, 5), torch.zeros(3, 4), 2)
    with pytest.raises(NotImplementedError):
        ns.group_points(torch.zeros(1, 3, 5), torch.zeros(1, 2, 2, dtype=torch.int32))
    meta = torch.device("meta")
    d, i = ns.knn(torch.empty(3, 120, device=meta), torch.empty(3, 1000, device=meta), 60)
    assert d.shape == (60, 1000) and i.shape == (60, 1000) and i.dtype == torch.int64
    d, i = ns.knn(torch.empty(4, 3, 120, device=meta), torch.empty(4, 3, 1000, device=meta), 60)
    assert d.shape == (4, 60, 1000) and d.dtype == torch.float32
    idx = ns.knn_graph(torch.empty(2, 64, 1000, device=meta), 10, 20)
    assert idx.shape == (2, 1000, 10) and idx.dtype == torch.int64   # the reference's dilation columns (M4:32)
    g = ns.group_points(torch.empty(2, 7, 100, device=meta), torch.empty(2, 50, 9, device=meta, dtype=torch.int32))
    assert g.shape == (2, 7, 50, 9)
    assert ns.group_points_grad(g, torch.empty(2, 50, 9, device=meta, dtype=torch.int32), 100).shape == (2, 7, 100)
