/* gcanet_b200 -- C-ABI of the B200-native DGCNN kNN-graph + EdgeConv path.
 *
 * One shared library (gcanet_b200/lib/libgcanet_b200.so, sm_100a only).  Every entry
 * point takes plain device pointers, integer sizes and a CUDA stream; nothing in this
 * header depends on PyTorch.  Rules that hold for every function:
 *
 *   - all data pointers are DEVICE pointers, fp32 unless the name says otherwise,
 *     dense row-major in the shape written in the comment;
 *   - the library never allocates, frees or synchronises: outputs and scratch
 *     (`ws`, sized by the matching *_workspace_bytes) are passed in by the caller,
 *     kernels are enqueued on `stream` and the call returns immediately, so every
 *     function may be captured in a CUDA graph (the one exception is the measurement
 *     probe gcanet_knn_probe_arm / _read, which says so where it is declared);
 *   - re-entrant, no mutable global state (per thread: the last error message and the
 *     probe's event pair), uses the calling thread's current device;
 *   - returns GCANET_OK (0) or a negative gcanet_status; gcanet_last_error() gives
 *     the message for the calling thread.  Nothing here ever exits the process (the
 *     reference's PN2 wrapper does: _ext-src/include/cuda_utils.h:30-39);
 *   - there is no CPU path: host pointers are a caller error.
 *
 * "Replaces" cites the reference interface each entry point stands in for; paths are
 * relative to the reference root, M4 = models/dgcnn-hais-concat-direct-4.py,
 * KNN = models/KNN_CUDA/knn_cuda, PN2 = models/Pointnet2_PyTorch-master/
 * pointnet2_ops_lib/pointnet2_ops.
 */
#ifndef GCANET_B200_H_
#define GCANET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCANET_ABI_VERSION 2

#if defined(__GNUC__)
#define GCANET_API __attribute__((visibility("default")))
#else
#define GCANET_API
#endif

typedef void *gcanet_stream_t; /* cudaStream_t */

typedef enum {
    GCANET_OK = 0,
    GCANET_ERR_INVALID_ARGUMENT = -1, /* bad size / null pointer / unsupported combination  */
    GCANET_ERR_WORKSPACE = -2,        /* ws too small or misaligned (needs 256-byte alignment) */
    GCANET_ERR_CUDA = -3,             /* a CUDA runtime call or launch failed                  */
    GCANET_ERR_UNSUPPORTED_DEVICE = -4 /* not an sm_100 device                                 */
} gcanet_status;

/* Neighbour metric of gcanet_knn_graph. */
typedef enum {
    GCANET_METRIC_L2 = 0,            /* -|x_i - x_j|^2 in expansion form        (M4:36-38)   */
    GCANET_METRIC_POINTS_NORMALS = 1 /* d_p * (1 + (2 - 2 n_i.n_j)), needs C = 6 (M4:61-73)   */
} gcanet_metric;

/* May be OR-ed into `metric`: run the plain brute-force CUDA-core scan even where an accelerated
 * path applies (tensor-core pruning for C = 64 / 128, spatial pruning for xyz clouds).  All
 * paths return the same fp32-exact neighbour lists; the flag exists for A/B tests and benchmarks. */
#define GCANET_KNN_FLAG_BRUTE_FORCE 0x100
#define GCANET_KNN_FLAG_NO_TENSOR_CORES GCANET_KNN_FLAG_BRUTE_FORCE
/* May be OR-ed into `metric` when k1 == k2: the caller only needs each point's neighbour SET, not
 * the nearest-first order (EdgeConv's max / sum over k is order-invariant).  The set is the same
 * exact set; paths that can save work by not ordering it (the tensor-core path re-ranks only the
 * candidates whose membership is in doubt) do so, the others ignore the flag. */
#define GCANET_KNN_FLAG_UNORDERED 0x200
/* May be OR-ed into `metric`: the tensor-core path scans every key tile instead of sorting the cloud
 * along its principal directions and skipping tiles whose bounding box cannot hold a neighbour.
 * Same result either way; for A/B tests and benchmarks. */
#define GCANET_KNN_FLAG_NO_PRUNE 0x400

/* Edge-feature variant of gcanet_graph_feature*. */
typedef enum {
    GCANET_EDGE_DIFF_CENTER = 0, /* (x_j - x_i, x_i)                      F = 2C  (M4:120-123) */
    GCANET_EDGE_NORMAL_ANGLE = 1 /* (clamp(n_i.n_j,+-.99), n_j-n_i, n_i)  F = 7, C = 6 (M4:189-204) */
} gcanet_edge_variant;

GCANET_API int gcanet_abi_version(void);
GCANET_API const char *gcanet_last_error(void);
GCANET_API const char *gcanet_status_string(int status);
/* Number of kernels this process has launched through the library so far (statistics only). */
GCANET_API unsigned long long gcanet_launch_count(void);
/* 0 when the current device is sm_100 (B200), GCANET_ERR_UNSUPPORTED_DEVICE otherwise. */
GCANET_API int gcanet_check_device(void);

/* ------------------------------------------------------------------ layouts
 * The reference keeps clouds channel-major, x[B][C][N] (train_new.py:25-26).  The fused
 * EdgeConv kernels work point-major, x[B][N][ld] with ld >= C (padding columns are
 * written as zero by cn_to_nc and ignored by nc_to_cn). */
GCANET_API int gcanet_cn_to_nc(const float *x_cn, float *x_nc, int B, int C, int N, int ld, gcanet_stream_t stream);
GCANET_API int gcanet_nc_to_cn(const float *x_nc, float *x_cn, int B, int C, int N, int ld, gcanet_stream_t stream);
/* x_nc = transpose(x_cn) + add_nc, add_nc [B][N][ld] laid out like x_nc (may alias x_nc): the sum of a channel-major
 * and a point-major incoming gradient in one pass (autograd's accumulation at x1/x2, M4:497-507, without the
 * intermediate tensor). */
GCANET_API int gcanet_cn_to_nc_add(const float *x_cn, const float *add_nc, float *x_nc, int B, int C, int N, int ld,
                                   gcanet_stream_t stream);

/* ------------------------------------------------------------------ kNN graph (torch path)
 * Replaces knn(x,k1,k2) M4:30-47, knn_points_normals(x,k1,k2) M4:50-90 and
 * splinenet.knn(x,k) models/splinenet.py:9-22: per cloud, the k2 nearest points of every
 * point under `metric`, nearest first, the point itself included, then the reference's
 * dilation sub-sampling columns 0, s, 2s, ... with s = k2 / k1 (integer division, M4:32).
 *   x      [B][C][N]
 *   idx64  [B][N][kout] int64 or NULL;  idx32 [B][N][kout] int32 or NULL  (at least one)
 *   kout = gcanet_knn_graph_columns(k1, k2);  requires 1 <= k1 <= k2 <= min(N, 1024).
 * Distances are fp32 with the reference's expansion arithmetic; no N x N matrix is
 * ever written to memory.  For C = 64 / 128 (L2 metric, k2 <= 128, N >= 128) candidates are
 * pruned on the tensor cores (tcgen05, bf16x3 split) and the survivors re-ranked in exact
 * fp32 -- for N >= 1024 after sorting the cloud along its three leading principal
 * directions, so that key tiles whose bounding box cannot hold a neighbour are never read
 * (GCANET_KNN_FLAG_NO_PRUNE scans every tile); xyz clouds (C = 3, L2, 1024 <= N < 32768,
 * k2 <= 128) take the same tensor-core scan with one K = 16 MMA per key tile (products of the
 * bf16x3 split and the norm in one step); the other xyz clouds (C = 3 outside those limits or
 * with GCANET_KNN_FLAG_NO_PRUNE, C = 6 points x normals; N >= 256) are sorted along a Hilbert
 * curve and scanned on the CUDA cores with AABB pruning; every other shape runs the
 * brute-force CUDA-core scan.  The result does not depend on the path. */
GCANET_API int gcanet_knn_graph_columns(int k1, int k2);
GCANET_API size_t gcanet_knn_graph_workspace_bytes(int B, int C, int N, int k2, int metric);
GCANET_API int gcanet_knn_graph(const float *x, int B, int C, int N, int k1, int k2, int metric,
                     int64_t *idx64, int32_t *idx32, void *ws, size_t ws_bytes,
                     gcanet_stream_t stream);

/* Measurement probe for the roofline line of bench.py -- not a data-path call and the only entry points that create
 * CUDA objects or wait for the device.  gcanet_knn_probe_arm(1) arms the calling thread: the next gcanet_knn_graph call
 * of that thread which takes a tensor-core scan brackets its distance-scan kernel(s) (knn_tcp_scan_kernel, plus the
 * full-scan launch for clouds without structure; NOT the preparation and NOT the exact re-rank) with a pair of CUDA
 * events on the call's stream, then disarms.  gcanet_knn_probe_read waits for that pair and stores the elapsed
 * milliseconds; it returns GCANET_ERR_INVALID_ARGUMENT when no scan has been bracketed since the thread was armed.
 * The events are created on the current device the first time a thread arms.  Never arm during stream capture. */
GCANET_API int gcanet_knn_probe_arm(int on);
GCANET_API int gcanet_knn_probe_read(float *scan_ms);

/* ------------------------------------------------------------------ kNN (KNN_CUDA path)
 * Replaces knn_device(ref,ref_nb,query,query_nb,dim,k,dist,ind,stream) KNN/csrc/cuda/knn.cpp:11-21
 * (kernels knn.cu:29-183) and the Python batch loop around it (KNN/__init__.py:41-74):
 * cross-set brute-force kNN with Euclidean (sqrt) distances, ascending, ties keep the
 * lower reference index first.
 *   ref [batch][dim][ref_nb], query [batch][dim][query_nb]
 *   dist [batch][k][query_nb], ind [batch][k][query_nb] int64, values in
 *   [index_base, ref_nb + index_base): index_base = 1 reproduces knn_device, 0 the
 *   Python-level knn() (KNN/__init__.py:43).  Unlike the reference no [ref_nb][query_nb]
 *   scratch matrix is needed (knn.cpp:36).  Requires 1 <= k <= min(ref_nb, 1024). */
GCANET_API size_t gcanet_knn_cuda_workspace_bytes(int batch, int dim, int ref_nb, int query_nb, int k);
GCANET_API int gcanet_knn_cuda(const float *ref, int ref_nb, const float *query, int query_nb, int dim, int k,
                    int batch, int index_base, float *dist, int64_t *ind, void *ws, size_t ws_bytes,
                    gcanet_stream_t stream);

/* ------------------------------------------------------------------ materialised edge features
 * Replaces the gather/repeat/cat part of get_graph_feature M4:103-123,
 * get_graph_feature_with_normals M4:140-160 and get_graph_feature_with_normals_g M4:177-204.
 *   x [B][C][N], idx [B][N][k] int64 (values in [0,N)), out [B][N][k][F]
 * (the reference returns this buffer viewed as [B][F][N][k] via permute(0,3,1,2)).
 * The gradient w.r.t. x (autograd of index/sub/cat/mul in the reference) is
 * gcanet_graph_feature_grad: grad_out [B][N][k][F] -> grad_x [B][C][N] (overwritten).
 * ws: a point-major staging copy, sized by the matching *_workspace_bytes.  An index
 * outside [0,N) is undefined behaviour, as in the reference. */
GCANET_API int gcanet_graph_feature_channels(int C, int variant);
GCANET_API size_t gcanet_graph_feature_workspace_bytes(int B, int C, int N, int k, int variant);
GCANET_API int gcanet_graph_feature(const float *x, const int64_t *idx, float *out, int B, int C, int N, int k,
                         int variant, void *ws, size_t ws_bytes, gcanet_stream_t stream);
GCANET_API size_t gcanet_graph_feature_grad_workspace_bytes(int B, int C, int N, int k, int variant);
GCANET_API int gcanet_graph_feature_grad(const float *grad_out, const float *x, const int64_t *idx, float *grad_x,
                              int B, int C, int N, int k, int variant, void *ws, size_t ws_bytes,
                              gcanet_stream_t stream);

/* ------------------------------------------------------------------ grouping (PN2 path)
 * Replaces group_points_kernel_wrapper / group_points_grad_kernel_wrapper
 * (PN2/_ext-src/src/group_points.cpp:4-10, kernels group_points_gpu.cu:8-28,43-64); same
 * argument order plus the stream.  points [b][c][n], idx [b][npoints][nsample] int32,
 * out [b][c][npoints][nsample]; grad_points [b][c][n] is overwritten (the reference
 * accumulates into a zeroed tensor, group_points.cpp:48-50). */
GCANET_API int gcanet_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                        const int32_t *idx, float *out, gcanet_stream_t stream);
GCANET_API int gcanet_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                             const int32_t *idx, float *grad_points, gcanet_stream_t stream);

/* ------------------------------------------------------------------ fused EdgeConv block
 * Replaces, for one layer, get_graph_feature(x, idx=idx) -> Conv2d(2C,Cout,1,bias=False) ->
 * GroupNorm(groups,Cout,eps) -> LeakyReLU(slope) -> max over k  (M4:469-481, M4:494-505) and
 * its autograd backward, without ever forming the [B][2C][N][k] edge tensor or the
 * [B][Cout][N][k] activation (see DESIGN.md for the identities used).
 *
 *   x_nc   [B][N][ldx]   point-major input, ldx >= C
 *   idx    [B][N][k]     int32 neighbour lists (gcanet_knn_graph's idx32)
 *   weight [Cout][2C]    conv weight, columns [0,C) multiply x_j - x_i, [C,2C) multiply x_i
 *   gamma, beta [Cout]   GroupNorm affine
 *   out_nc [B][N][Cout]  and, if not NULL, out_cn [B][Cout][N] (the reference layout)
 *   saved  opaque, gcanet_edgeconv_saved_bytes(); must be kept untouched until backward
 *   ws     scratch, gcanet_edgeconv_workspace_bytes()
 * Constraints: Cout % 32 == 0, Cout <= 256, Cout % groups == 0, (Cout/groups) % (Cout/32) == 0,
 * C <= 256, 1 <= k <= 255, N * 2 * Cout < 2^30 per cloud (rows are addressed with 32-bit offsets).
 *
 * backward: grad_out_nc [B][N][Cout] -> grad_x_nc [B][N][ldx] (or NULL to skip; padding
 * columns are zeroed), grad_weight [Cout][2C], grad_gamma, grad_beta [Cout], all overwritten. */
typedef struct {
    int B, N, C, ldx, Cout, k, groups;
    float eps;   /* GroupNorm epsilon, reference 1e-5 */
    float slope; /* LeakyReLU negative slope, reference 0.2 */
    /* 0: the projected operand [P|Q] is kept in fp32 (parity mode, default).  1: it is rounded to bf16 where the projection
     * GEMM writes it ("bf16 activations" of BASELINE configs[1]): the gather reads half the bytes, the saved tensor is half
     * the size; accumulation, GroupNorm statistics and every gradient stay fp32.  Outputs then agree with the reference
     * to bf16 rounding of the pre-norm activation (2^-9 relative), see tests/test_gpu_bf16.py. */
    int storage_bf16;
} gcanet_edgeconv_desc;

GCANET_API size_t gcanet_edgeconv_saved_bytes(const gcanet_edgeconv_desc *d);
GCANET_API size_t gcanet_edgeconv_workspace_bytes(const gcanet_edgeconv_desc *d);
GCANET_API int gcanet_edgeconv_forward(const gcanet_edgeconv_desc *d, const float *x_nc, const int32_t *idx,
                            const float *weight, const float *gamma, const float *beta,
                            float *out_nc, float *out_cn, void *saved, void *ws, size_t ws_bytes,
                            gcanet_stream_t stream);
GCANET_API int gcanet_edgeconv_backward(const gcanet_edgeconv_desc *d, const float *x_nc, const int32_t *idx,
                             const float *weight, const float *gamma, const float *beta,
                             const float *grad_out_nc, const void *saved, float *grad_x_nc,
                             float *grad_weight, float *grad_gamma, float *grad_beta, void *ws,
                             size_t ws_bytes, gcanet_stream_t stream);

/* ------------------------------------------------------------------ fused EdgeConv on normals
 * Replaces get_graph_feature_with_normals_g(points, idx=idx) -> conv_normal = Conv2d(7,Cout,1,bias=False)
 * -> GroupNorm -> LeakyReLU -> max over k (M4:584-587, M4:691-693) and the parameter gradients of its
 * backward.  The 7-channel edge feature (clamp(n_i.n_j, +-.99), n_j - n_i, n_i) is rebuilt per edge
 * from the neighbour's normal; neither [B][7][N][k] nor [B][Cout][N][k] is formed.
 *   x_nc   [B][N][ldx], ldx >= 6, normals in columns 3..5 (gcanet_cn_to_nc of the reference's [B][6][N])
 *   idx    [B][N][k] int32;  weight [Cout][7];  gamma, beta [Cout];  Cout in {32, 64, 128}
 *   out_nc [B][N][Cout], out_cn [B][Cout][N] or NULL
 * backward produces grad_weight [Cout][7], grad_gamma, grad_beta (all overwritten).  There is no
 * gradient w.r.t. x: in the reference this head is fed the input points and normals (data). */
typedef struct {
    int B, N, ldx, Cout, k, groups;
    float eps, slope;
} gcanet_normal_edge_desc;

GCANET_API size_t gcanet_normal_edgeconv_saved_bytes(const gcanet_normal_edge_desc *d);
GCANET_API size_t gcanet_normal_edgeconv_workspace_bytes(const gcanet_normal_edge_desc *d);
GCANET_API int gcanet_normal_edgeconv_forward(const gcanet_normal_edge_desc *d, const float *x_nc, const int32_t *idx,
                                              const float *weight, const float *gamma, const float *beta,
                                              float *out_nc, float *out_cn, void *saved, void *ws, size_t ws_bytes,
                                              gcanet_stream_t stream);
GCANET_API int gcanet_normal_edgeconv_backward(const gcanet_normal_edge_desc *d, const float *x_nc, const int32_t *idx,
                                               const float *weight, const float *gamma, const float *beta,
                                               const float *grad_out_nc, const void *saved, float *grad_weight,
                                               float *grad_gamma, float *grad_beta, void *ws, size_t ws_bytes,
                                               gcanet_stream_t stream);

/* ------------------------------------------------------------------ encoder tail (global feature)
 * Replaces  x = relu(bnmlp1(mlp1(cat(x1, x2, x3))));  x4 = x.max(dim=2)[0]   (M4:507-510: Conv1d(256 -> 1024, 1, bias) ->
 * GroupNorm(8, 1024) -> ReLU -> max over the N points) and its autograd backward, without forming the [B][Cout][N]
 * activation (the reference keeps three copies of it).  The broadcast + concat of M4:510-511 into [B][1280][N] is left
 * to the caller: 1024 of those channels are one value per cloud (fold them into the next layer's bias).
 *   x_nc   [B][N][K]   point-major concatenation of x1 | x2 | x3 (K = 256), 16-byte aligned
 *   weight [Cout][K]   mlp1.weight viewed as a matrix;  bias [Cout] or NULL;  gamma, beta [Cout]
 *   out    [B][Cout]   x4
 * backward: grad_out [B][Cout] -> grad_x_nc [B][N][K] (or NULL to skip), grad_weight [Cout][K], grad_bias [Cout] (or NULL),
 * grad_gamma, grad_beta [Cout]; all overwritten.  Constraints: K = 256, Cout % 128 == 0, Cout % groups == 0. */
typedef struct {
    int B, N, K, Cout, groups;
    float eps;
} gcanet_global_feature_desc;

GCANET_API size_t gcanet_global_feature_saved_bytes(const gcanet_global_feature_desc *d);
GCANET_API size_t gcanet_global_feature_workspace_bytes(const gcanet_global_feature_desc *d);
GCANET_API int gcanet_global_feature_forward(const gcanet_global_feature_desc *d, const float *x_nc, const float *weight,
                                             const float *bias, const float *gamma, const float *beta, float *out, void *saved,
                                             void *ws, size_t ws_bytes, gcanet_stream_t stream);
GCANET_API int gcanet_global_feature_backward(const gcanet_global_feature_desc *d, const float *x_nc, const float *weight,
                                              const float *bias, const float *gamma, const float *beta, const float *grad_out,
                                              const void *saved, float *grad_x_nc, float *grad_weight, float *grad_bias,
                                              float *grad_gamma, float *grad_beta, void *ws, size_t ws_bytes,
                                              gcanet_stream_t stream);

/* ------------------------------------------------------------------ GroupNorm + ReLU of the per-point heads
 * Replaces F.relu(GroupNorm(x)) on channel-major activations -- F.relu(self.bn1(self.conv1(x))) and its siblings,
 * M4:644-645, 650, 661, 698, 713 -- and its autograd backward.  A group's values are contiguous in [B][C][N], so the
 * statistics are split over many CTAs (fp64 partials, fixed-order finalize: deterministic).
 *   x, y, grad_y, grad_x  [B][C][N];  gamma, beta, grad_gamma, grad_beta [C];  stats [B][groups][2] (mean, rstd): written by
 *   forward, read by backward (the only saved state; y is recomputed from x where the ReLU mask is needed)
 *   act: 0 = none, 1 = ReLU.   y may alias x (in place) in forward only when backward is not needed. */
typedef struct {
    int B, C, N, groups;
    float eps;
    int act;
} gcanet_group_norm_desc;

GCANET_API size_t gcanet_group_norm_workspace_bytes(const gcanet_group_norm_desc *d);
GCANET_API int gcanet_group_norm_forward(const gcanet_group_norm_desc *d, const float *x, const float *gamma, const float *beta,
                                         float *y, float *stats, void *ws, size_t ws_bytes, gcanet_stream_t stream);
GCANET_API int gcanet_group_norm_backward(const gcanet_group_norm_desc *d, const float *x, const float *gamma, const float *beta,
                                          const float *stats, const float *grad_y, float *grad_x, float *grad_gamma,
                                          float *grad_beta, void *ws, size_t ws_bytes, gcanet_stream_t stream);

/* ------------------------------------------------------------------ offset-prediction block
 * Replaces OFFSET_PRED_MODULE.forward with KPAM and cos_dist (M4:326-452) and its autograd backward: per cloud S key
 * points (the caller passes their indices; the reference re-seeds numpy with 1234 and takes the first S of a shuffle,
 * M4:403-406), cosine similarity of every point's instance feature to the keys', the k most similar keys (descending,
 * ties by key index), attention softmax_k(W2 relu(W1 d)) on the similarity values, edge feature
 * a_ik (f_j ; p_j - p_i) -> Conv2d(131 -> 128, no bias) -> GroupNorm -> LeakyReLU -> max over k, concatenated with the
 * point's own feature, Conv1d(256 -> 3).  Nothing of size N x S x C or N x k x C is stored.
 *   points [B][N][3], feature [B][N][128], inst [B][N][E] (point-major, as the module receives them)
 *   key_index [S] int32;  conv_w [128][131];  gamma, beta [128];  att_w1, att_w2 [k][k];  off_w [3][256];  off_b [3]
 *   out [B][3][N]
 * backward: grad_out [B][3][N] -> grad_feature [B][N][128], grad_inst [B][N][E] and all parameter gradients (overwritten);
 * points are data.  Constraints: S % 4 == 0, S <= 128, k <= min(32, S), E % 4 == 0, E <= 256 (the key tables of a cloud
 * must fit in shared memory: S * (E + 129) floats). */
typedef struct {
    int B, N, S, k, E, groups;
    float eps, slope;
} gcanet_offset_desc;

GCANET_API size_t gcanet_offset_pred_saved_bytes(const gcanet_offset_desc *d);
GCANET_API size_t gcanet_offset_pred_workspace_bytes(const gcanet_offset_desc *d);
GCANET_API int gcanet_offset_pred_forward(const gcanet_offset_desc *d, const float *points, const float *feature, const float *inst,
                                          const int32_t *key_index, const float *conv_w, const float *gamma, const float *beta,
                                          const float *att_w1, const float *att_w2, const float *off_w, const float *off_b,
                                          float *out, void *saved, void *ws, size_t ws_bytes, gcanet_stream_t stream);
GCANET_API int gcanet_offset_pred_backward(const gcanet_offset_desc *d, const float *points, const float *feature, const float *inst,
                                           const int32_t *key_index, const float *conv_w, const float *gamma, const float *beta,
                                           const float *att_w1, const float *att_w2, const float *off_w, const float *off_b,
                                           const float *grad_out, const void *saved, float *grad_feature, float *grad_inst,
                                           float *grad_conv_w, float *grad_gamma, float *grad_beta, float *grad_att_w1,
                                           float *grad_att_w2, float *grad_off_w, float *grad_off_b, void *ws, size_t ws_bytes,
                                           gcanet_stream_t stream);

/* ------------------------------------------------------------------ dense affinity + gated ball query (grouping front end)
 * compute_batch_adjacency_matrix (M4:210-233) and ballquery_batch_p (softgroup/ops/src/bfs_cluster/bfs_cluster.cpp:20-46,
 * kernel bfs_cluster.cu:18-77; Python loop softgroup/ops/functions.py:460-472).  All pointers are device pointers.
 *
 * gcanet_affinity_matrix        adj [n][n] = exp(-(d_ik / d_max)^2 / (2 sigma^2)), zero diagonal, d = Euclidean distance of the
 *                               rows of x [n][C], d_max its largest value (the reference's global min is the zero diagonal).
 * gcanet_ball_query_dense       the reference kernel's contract on the dense matrices it is given: per point the neighbours k of
 *                               its own batch segment with |p_i - p_k|^2 < radius^2, adj_inst[i][k] > thr_inst, adj_para[i][k] >
 *                               thr_para; ascending k, at most 3000; start_len [n][2] = (start, length), lists in point order
 *                               (the reference: atomicAdd order); *total = neighbours found (re-run with more capacity if larger).
 * gcanet_affinity_ball_query    the three calls fused, no n x n matrix: one pass per feature space for d_max per segment, then the
 *                               radius search evaluates feature distances only for spatial neighbours.
 * gcanet_pairwise_max_distance  dmax2 [segments] = largest squared pairwise distance inside each segment. */
GCANET_API size_t gcanet_affinity_workspace_bytes(int segments);
GCANET_API int gcanet_pairwise_max_distance(const float *x, const int32_t *seg, int n, int C, int segments, int max_segment,
                                            float *dmax2, gcanet_stream_t stream);
GCANET_API int gcanet_affinity_matrix(const float *x, int n, int C, float sigma, float *adj, void *ws, size_t ws_bytes,
                                      gcanet_stream_t stream);
GCANET_API int gcanet_ball_query_dense(const float *xyz, const int32_t *batch_offsets, int n, int segments, const float *adj_inst,
                                       float thr_inst, const float *adj_para, float thr_para, float radius, int32_t *idx,
                                       long long capacity, int32_t *start_len, long long *total, gcanet_stream_t stream);
GCANET_API int gcanet_affinity_ball_query(const float *xyz, const int32_t *batch_offsets, int n, int segments, int max_segment,
                                          const float *f_inst, int Ci, float thr_inst, const float *f_para, int Cp, float thr_para,
                                          float sigma, float radius, int32_t *idx, long long capacity, int32_t *start_len,
                                          long long *total, void *ws, size_t ws_bytes, gcanet_stream_t stream);

/* ------------------------------------------------------------------ input side: batch preparation on the device
 * Replaces, for a batch of raw samples resident in device memory, what ABCDataset.__getitem__ does after reading a sample
 * (dataloader/ABCDataset_new.py:77-141), getInstanceInfo (:157-178) and collate_fn's stacking (:182-295): small instances
 * (<= min_points raw points) become background, kept ones are renumbered in order of first appearance, the subsample is
 * gathered, per-instance centroids give the offset labels.
 *   points, normals [B][n_raw][3];  labels, prim [B][n_raw] int32 (labels in [0, max_labels));  t_param [B][n_raw][22]
 *   sub_index [B][n_sub] int32: the subsample (the reference draws it with np.random.choice(..., replace=False), :120)
 *   cloud_cn [B][6][n_sub] (reference layout, xyz then normals), cloud_nc [B][n_sub][8] (xyz | normal | 0 0, point-major)
 *   i_gt, t_gt, i_gt_clean [B][n_sub] int32;  t_param_out [B][n_sub][22];  pt_offset_label [B][n_sub][3]
 *   inst_num [B];  inst_pointnum, inst_cls [B][max_labels] (first inst_num[b] entries valid);  status [B]: 0 ok,
 *   1 = a label outside [0, max_labels) (nothing else is written for that cloud). */
typedef struct {
    int B, n_raw, n_sub, max_labels, min_points, num_primitives;
} gcanet_prepare_desc;

GCANET_API int gcanet_prepare_samples(const gcanet_prepare_desc *d, const float *points, const float *normals, const int32_t *labels,
                                      const int32_t *prim, const float *t_param, const int32_t *sub_index, float *cloud_cn,
                                      float *cloud_nc, int32_t *i_gt, int32_t *t_gt, int32_t *i_gt_clean, float *t_param_out,
                                      float *pt_offset_label, int32_t *inst_num, int32_t *inst_pointnum, int32_t *inst_cls,
                                      int32_t *status, gcanet_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GCANET_B200_H_ */
