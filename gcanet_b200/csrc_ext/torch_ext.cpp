// PyTorch C++ extension over the C-ABI of libgcanet_b200.so.
//
// The reference binds its native kernels as pybind11 modules built by torch's cpp_extension:
//   knn.knn(ref, query, k)                               models/KNN_CUDA/knn_cuda/csrc/cuda/knn.cpp:23-61
//   _ext.group_points(points, idx)                       PN2 _ext-src/src/group_points.cpp:12-36, bindings.cpp:17
//   _ext.group_points_grad(grad_out, idx, n)             PN2 _ext-src/src/group_points.cpp:38-62, bindings.cpp:18
// This file is their counterpart: the same three functions (names, argument order, dtypes, shapes, 1-based indices of
// knn, "CPU not supported"), plus knn_graph for the torch path (M4:30-90), each a few lines of checking and allocation
// around ONE call into the PyTorch-free library (include/gcanet_b200.h).  The same functions are registered with the
// dispatcher as torch.ops.gcanet_b200_native.* (CUDA key only: a CPU tensor finds no kernel, there is no fallback).
//
// Built ahead of time by gcanet_b200/build.py (g++, no nvcc needed) into gcanet_b200/lib/gcanet_b200_ext.so.
#include <torch/extension.h>
#include <torch/library.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include <tuple>
#include <vector>

#include "gcanet_b200.h"

namespace {

void check_ok(int rc, const char *what) {
    TORCH_CHECK(rc == GCANET_OK, "gcanet_b200 ", what, ": ", gcanet_last_error(), " (status ", rc, ")");
}

at::Tensor byte_workspace(size_t bytes, const at::Tensor &like) {
    return at::empty({(int64_t)(bytes < 256 ? 256 : bytes)}, like.options().dtype(at::kByte));
}

void check_float_cuda(const at::Tensor &t, const char *name) {
    TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
    TORCH_CHECK(t.scalar_type() == at::kFloat, name, " must be at::kFloat");
    TORCH_CHECK(t.is_cuda(), name, " must be on CUDA");
}

// ref [dim][Nr] and query [dim][Nq] (the reference's shapes), or batched [B][dim][Nr] / [B][dim][Nq] in ONE launch.
// Returns (dist [k][Nq] Euclidean ascending, ind [k][Nq] int64 in [index_base, Nr + index_base)); index_base = 1 is
// what knn.cpp:23-56 returns (knn.cu:119), 0 what the Python-level knn() makes of it (knn_cuda/__init__.py:43).
std::tuple<at::Tensor, at::Tensor> knn_impl(const at::Tensor &ref, const at::Tensor &query, int64_t k, int64_t index_base) {
    check_float_cuda(ref, "ref");
    check_float_cuda(query, "query");
    TORCH_CHECK(ref.dim() == query.dim() && (ref.dim() == 2 || ref.dim() == 3), "ref / query must both be [dim, n] or [B, dim, n]");
    TORCH_CHECK(ref.get_device() == query.get_device(), "ref and query must be on the same device");
    const bool batched = ref.dim() == 3;
    const int64_t B = batched ? ref.size(0) : 1;
    TORCH_CHECK(!batched || query.size(0) == B, "ref and query must hold the same number of clouds");
    const int64_t dim = ref.size(-2), ref_nb = ref.size(-1), query_nb = query.size(-1);
    TORCH_CHECK(query.size(-2) == dim, "ref and query must have the same dimension");
    c10::cuda::CUDAGuard guard(ref.device());
    std::vector<int64_t> shape = batched ? std::vector<int64_t>{B, k, query_nb} : std::vector<int64_t>{k, query_nb};
    at::Tensor dist = at::empty(shape, query.options());                     // no [Nr][Nq] scratch (knn.cpp:36)
    at::Tensor ind = at::empty(shape, query.options().dtype(at::kLong));
    at::Tensor ws = byte_workspace(gcanet_knn_cuda_workspace_bytes((int)B, (int)dim, (int)ref_nb, (int)query_nb, (int)k), query);
    check_ok(gcanet_knn_cuda(ref.data_ptr<float>(), (int)ref_nb, query.data_ptr<float>(), (int)query_nb, (int)dim, (int)k, (int)B,
                             (int)index_base, dist.data_ptr<float>(), ind.data_ptr<int64_t>(), ws.data_ptr(), (size_t)ws.numel(),
                             at::cuda::getCurrentCUDAStream().stream()),
             "knn");
    return {dist, ind};
}

std::vector<at::Tensor> knn(at::Tensor &ref, at::Tensor &query, const int k) {
    auto r = knn_impl(ref, query, k, 1);
    return {std::get<0>(r), std::get<1>(r)};
}

// x [B][C][N] -> idx [B][N][kout] int64, nearest first, self included, the reference's dilation columns (M4:30-47);
// metric 1 = points x normals (M4:50-90)
at::Tensor knn_graph(const at::Tensor &x, int64_t k1, int64_t k2, int64_t metric) {
    check_float_cuda(x, "x");
    TORCH_CHECK(x.dim() == 3, "x must be [B, C, N]");
    const int B = (int)x.size(0), C = (int)x.size(1), N = (int)x.size(2);
    const int kout = gcanet_knn_graph_columns((int)k1, (int)k2);
    TORCH_CHECK(kout > 0, "need 1 <= k1 <= k2 (k1=", k1, ", k2=", k2, ")");
    c10::cuda::CUDAGuard guard(x.device());
    at::Tensor idx = at::empty({B, N, kout}, x.options().dtype(at::kLong));
    at::Tensor ws = byte_workspace(gcanet_knn_graph_workspace_bytes(B, C, N, (int)k2, (int)metric), x);
    check_ok(gcanet_knn_graph(x.data_ptr<float>(), B, C, N, (int)k1, (int)k2, (int)metric, idx.data_ptr<int64_t>(), nullptr,
                              ws.data_ptr(), (size_t)ws.numel(), at::cuda::getCurrentCUDAStream().stream()),
             "knn_graph");
    return idx;
}

void check_group_args(const at::Tensor &data, const at::Tensor &idx, const char *data_name) {
    TORCH_CHECK(data.is_contiguous(), data_name, " must be a contiguous tensor");
    TORCH_CHECK(idx.is_contiguous(), "idx must be a contiguous tensor");
    TORCH_CHECK(data.scalar_type() == at::kFloat, data_name, " must be a float tensor");
    TORCH_CHECK(idx.scalar_type() == at::kInt, "idx must be an int tensor");
    TORCH_CHECK(data.is_cuda(), "CPU not supported");                          // group_points.cpp:32,58
    TORCH_CHECK(idx.is_cuda(), "idx must be a CUDA tensor");
    TORCH_CHECK(idx.dim() == 3, "idx must be [B, npoint, nsample]");
}

// points [B][C][N], idx [B][npoint][nsample] int32 -> [B][C][npoint][nsample]
at::Tensor group_points(at::Tensor points, at::Tensor idx) {
    check_group_args(points, idx, "points");
    TORCH_CHECK(points.dim() == 3 && idx.size(0) == points.size(0), "points must be [B, C, N] with idx's batch size");
    c10::cuda::CUDAGuard guard(points.device());
    at::Tensor out = at::empty({points.size(0), points.size(1), idx.size(1), idx.size(2)}, points.options());   // every element is written
    check_ok(gcanet_group_points((int)points.size(0), (int)points.size(1), (int)points.size(2), (int)idx.size(1), (int)idx.size(2),
                                 points.data_ptr<float>(), idx.data_ptr<int>(), out.data_ptr<float>(),
                                 at::cuda::getCurrentCUDAStream().stream()),
             "group_points");
    return out;
}

// grad_out [B][C][npoint][nsample], idx as above -> grad_points [B][C][n]
at::Tensor group_points_grad(at::Tensor grad_out, at::Tensor idx, const int n) {
    check_group_args(grad_out, idx, "grad_out");
    TORCH_CHECK(grad_out.dim() == 4 && idx.size(0) == grad_out.size(0) && idx.size(1) == grad_out.size(2) &&
                    idx.size(2) == grad_out.size(3),
                "grad_out must be [B, C, npoint, nsample] matching idx");
    c10::cuda::CUDAGuard guard(grad_out.device());
    at::Tensor out = at::zeros({grad_out.size(0), grad_out.size(1), (int64_t)n}, grad_out.options());           // accumulated into
    check_ok(gcanet_group_points_grad((int)grad_out.size(0), (int)grad_out.size(1), n, (int)idx.size(1), (int)idx.size(2),
                                      grad_out.data_ptr<float>(), idx.data_ptr<int>(), out.data_ptr<float>(),
                                      at::cuda::getCurrentCUDAStream().stream()),
             "group_points_grad");
    return out;
}

at::Tensor group_points_grad_op(const at::Tensor &grad_out, const at::Tensor &idx, int64_t n) { return group_points_grad(grad_out, idx, (int)n); }
at::Tensor group_points_op(const at::Tensor &points, const at::Tensor &idx) { return group_points(points, idx); }

}  // namespace

TORCH_LIBRARY(gcanet_b200_native, m) {
    m.def("knn(Tensor ref, Tensor query, int k, int index_base=1) -> (Tensor, Tensor)");
    m.def("knn_graph(Tensor x, int k1, int k2, int metric=0) -> Tensor");
    m.def("group_points(Tensor points, Tensor idx) -> Tensor");
    m.def("group_points_grad(Tensor grad_out, Tensor idx, int n) -> Tensor");
}

TORCH_LIBRARY_IMPL(gcanet_b200_native, CUDA, m) {
    m.impl("knn", &knn_impl);
    m.impl("knn_graph", &knn_graph);
    m.impl("group_points", &group_points_op);
    m.impl("group_points_grad", &group_points_grad_op);
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "gcanet_b200: PyTorch C++ extension over the C-ABI of libgcanet_b200.so";
    m.def("knn", &knn, "KNN_CUDA's knn(ref, query, k) -> [dist, ind] (1-based), one launch, no scratch matrix", py::arg("ref"),
          py::arg("query"), py::arg("k"));
    m.def("knn_graph", &knn_graph, "knn(x, k1, k2) / knn_points_normals of the DGCNN backbone -> idx int64", py::arg("x"),
          py::arg("k1"), py::arg("k2"), py::arg("metric") = 0);
    m.def("group_points", &group_points, "PN2 group_points(points, idx)", py::arg("points"), py::arg("idx"));
    m.def("group_points_grad", &group_points_grad, "PN2 group_points_grad(grad_out, idx, n)", py::arg("grad_out"), py::arg("idx"),
          py::arg("n"));
    m.def("abi_version", []() { return gcanet_abi_version(); });
}
