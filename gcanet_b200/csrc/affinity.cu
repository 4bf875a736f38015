// Dense-affinity front end of GCANet's proposal grouping:
//   compute_batch_adjacency_matrix  M4:210-233   (torch.cdist -> global min/max normalisation -> exp(-d^2 / 2 sigma^2), zero diagonal)
//   ballquery_batch_p_cuda_         softgroup/ops/src/bfs_cluster/bfs_cluster.cu:18-77  (radius search gated by TWO dense n x n
//                                   affinity matrices, 3000-int array per thread, list order = atomicAdd order)
//
// Per (cloud, class) the reference builds two n x n fp32 matrices (400 MB each at n = 10^4) to answer, per point, "which points
// within radius r have affinity above a threshold".  Because the matrix is normalised by its own maximum (its minimum is the
// zeroed diagonal), the gate is a plain distance test:  exp(-(d / d_max)^2 / 2 sigma^2) > t  <=>  d < d_max sigma sqrt(-2 ln t).
// So the fused path needs one number per feature space (the largest pairwise distance, one tiled pass) and then a radius search
// that evaluates feature distances only for the few pairs that are spatial neighbours: no n x n matrix, no per-thread array,
// and neighbour lists in a deterministic order (point order, ascending neighbour index inside a list -- the reference's inner
// order; its outer order is whatever atomicAdd produced).
//
// Also here, for API parity: the dense matrix builder and the reference-signature ball query that consumes dense matrices.
#include "common.cuh"

#include <math.h>
#include <math_constants.h>

namespace gcanet {

constexpr unsigned AFULL = 0xffffffffu;
constexpr int AF_CAP = 3000;          // neighbours kept per point (idx_temp[3000], bfs_cluster.cu:30,53-57)

// ---------------------------------------------------------------------------------------------- largest pairwise distance
// dmax2[s] = max over pairs (i, k) of segment s of |x_i - x_k|^2 (exact differences, fp32).  64 x 64 pair tiles, upper
// triangle only; block (16, 16), 4 x 4 pairs per thread, C in chunks of 16 through shared memory.
__global__ void __launch_bounds__(256) af_dmax_kernel(const float *__restrict__ x, const int *__restrict__ seg, int C,
                                                      unsigned *__restrict__ dmax2_bits) {
    __shared__ float sa[16][65], sb[16][65];
    __shared__ float red[8];
    const int s = blockIdx.z;
    const int lo = seg[s], n = seg[s + 1] - lo;
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj < bi || bi * 64 >= n || bj * 64 >= n) return;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int c0 = 0; c0 < C; c0 += 16) {
        for (int e = threadIdx.x; e < 64 * 16; e += 256) {
            const int r = e >> 4, c = e & 15;
            const int ia = bi * 64 + r, ib = bj * 64 + r;
            sa[c][r] = (ia < n && c0 + c < C) ? x[(size_t)(lo + ia) * C + c0 + c] : 0.f;
            sb[c][r] = (ib < n && c0 + c < C) ? x[(size_t)(lo + ib) * C + c0 + c] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            float av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { av[i] = sa[c][ty * 4 + i]; bv[i] = sb[c][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) { const float t = av[i] - bv[j]; acc[i][j] = fmaf(t, t, acc[i][j]); }
        }
        __syncthreads();
    }
    float m = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (bi * 64 + ty * 4 + i < n && bj * 64 + tx * 4 + j < n) m = fmaxf(m, acc[i][j]);
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(AFULL, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
        atomicMax(dmax2_bits + s, __float_as_uint(m));           // non-negative floats order like their bit patterns
    }
}

// ---------------------------------------------------------------------------------------------- dense matrix (API parity)
// adj[i][k] = exp(-(d_ik / d_max)^2 / (2 sigma^2)), adj[i][i] = 0      one 64 x 64 tile per CTA, single segment
__global__ void __launch_bounds__(256) af_matrix_kernel(const float *__restrict__ x, int n, int C, const float *__restrict__ dmax2,
                                                        float inv_two_sigma2, float *__restrict__ adj) {
    __shared__ float sa[16][65], sb[16][65];
    const int bi = blockIdx.y, bj = blockIdx.x;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int c0 = 0; c0 < C; c0 += 16) {
        for (int e = threadIdx.x; e < 64 * 16; e += 256) {
            const int r = e >> 4, c = e & 15;
            const int ia = bi * 64 + r, ib = bj * 64 + r;
            sa[c][r] = (ia < n && c0 + c < C) ? x[(size_t)ia * C + c0 + c] : 0.f;
            sb[c][r] = (ib < n && c0 + c < C) ? x[(size_t)ib * C + c0 + c] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            float av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { av[i] = sa[c][ty * 4 + i]; bv[i] = sb[c][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) { const float t = av[i] - bv[j]; acc[i][j] = fmaf(t, t, acc[i][j]); }
        }
        __syncthreads();
    }
    const float dm2 = *dmax2;
    const float inv = dm2 > 0.f ? 1.f / dm2 : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = bi * 64 + ty * 4 + i;
        if (r >= n) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = bj * 64 + tx * 4 + j;
            if (c < n) adj[(size_t)r * n + c] = r == c ? 0.f : expf(-(acc[i][j] * inv) * inv_two_sigma2);
        }
    }
}

// ---------------------------------------------------------------------------------------------- neighbour lists
// One warp per point; lanes walk the segment 32 candidates at a time, a ballot keeps the list in ascending neighbour order.
// MODE 0: count only (start_len[i][1]); MODE 1: write idx at start_len[i][0].
struct AfQueryArgs {
    const float *xyz;                 // [n][3]
    const int *seg;                   // [S + 1]
    // fused path
    const float *f_inst, *f_para;     // [n][Ci], [n][Cp]
    int Ci, Cp;
    const float *dmax2_inst, *dmax2_para;   // [S]
    float c2_inst, c2_para;           // squared threshold factors: pair passes iff d2 < dmax2 * c2 (c2 < 0: never, +inf: always)
    // dense path
    const float *adj_inst, *adj_para; // [n][n] (single segment) or null
    float thr_inst, thr_para;
    float r2;
    int n, S;
    int include_self;                 // both thresholds negative: the zero diagonal passes too (reference semantics)
    int *start_len;                   // [n][2]
    int *idx;                         // [capacity]
    long long capacity;
};

__device__ __forceinline__ bool af_gate(const float *f, int C, size_t i, size_t k, float lim2) {
    if (lim2 < 0.f) return false;                         // threshold >= 1: exp(.) <= 1 never exceeds it
    if (lim2 == CUDART_INF_F) return true;                // threshold <= 0: every off-diagonal entry is positive
    float d2 = 0.f;
    for (int c = 0; c < C; ++c) { const float t = f[i * C + c] - f[k * C + c]; d2 = fmaf(t, t, d2); }
    return d2 < lim2;
}

template <int MODE, bool DENSE>
__global__ void __launch_bounds__(256) af_query_kernel(AfQueryArgs a) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= a.n) return;
    const int i = warp;
    // segment of point i (S is small: linear search)
    int s = 0;
    while (s + 1 < a.S && a.seg[s + 1] <= i) ++s;
    const int lo = a.seg[s], hi = a.seg[s + 1];
    const float ox = a.xyz[(size_t)i * 3], oy = a.xyz[(size_t)i * 3 + 1], oz = a.xyz[(size_t)i * 3 + 2];
    float lim_i = 0.f, lim_p = 0.f;
    if (!DENSE) {
        lim_i = a.c2_inst < 0.f ? -1.f : (a.c2_inst == CUDART_INF_F ? CUDART_INF_F : a.dmax2_inst[s] * a.c2_inst);
        lim_p = a.c2_para < 0.f ? -1.f : (a.c2_para == CUDART_INF_F ? CUDART_INF_F : a.dmax2_para[s] * a.c2_para);
    }
    int cnt = 0;
    const long long start = MODE == 1 ? a.start_len[(size_t)i * 2] : 0;
    for (int k0 = lo; k0 < hi && cnt < AF_CAP; k0 += 32) {
        const int k = k0 + lane;
        bool pass = false;
        if (k < hi) {
            const float dx = ox - a.xyz[(size_t)k * 3], dy = oy - a.xyz[(size_t)k * 3 + 1], dz = oz - a.xyz[(size_t)k * 3 + 2];
            // the reference's expression order (bfs_cluster.cu:45-46), no contraction
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            if (d2 < a.r2) {
                if (DENSE) {
                    pass = a.adj_inst[(size_t)i * a.n + k] > a.thr_inst && a.adj_para[(size_t)i * a.n + k] > a.thr_para;
                } else {
                    pass = (k != i || a.include_self) && af_gate(a.f_inst, a.Ci, i, k, lim_i) && af_gate(a.f_para, a.Cp, i, k, lim_p);
                }
            }
        }
        const unsigned m = __ballot_sync(AFULL, pass);
        const int pos = cnt + __popc(m & ((1u << lane) - 1));
        if (MODE == 1 && pass && pos < AF_CAP && start + pos < a.capacity) a.idx[start + pos] = k;
        cnt += __popc(m);
    }
    if (cnt > AF_CAP) cnt = AF_CAP;
    if (MODE == 0 && lane == 0) a.start_len[(size_t)i * 2 + 1] = cnt;
}

// start_len[i][0] = exclusive prefix sum of the counts; total[0] = their sum.  One CTA (n <= a few 10^5).
__global__ void __launch_bounds__(1024) af_scan_kernel(int *__restrict__ start_len, int n, long long *__restrict__ total) {
    __shared__ long long part[1024];
    const int t = threadIdx.x;
    const int per = (n + 1023) / 1024, lo = t * per, hi = min(n, lo + per);
    long long s = 0;
    for (int i = lo; i < hi; ++i) s += start_len[(size_t)i * 2 + 1];
    part[t] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const long long v = t >= o ? part[t - o] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    long long run = t ? part[t - 1] : 0;
    for (int i = lo; i < hi; ++i) {
        start_len[(size_t)i * 2] = (int)(run > 2147483647ll ? 2147483647ll : run);
        run += start_len[(size_t)i * 2 + 1];
    }
    if (t == 1023) total[0] = part[1023];
}

__global__ void af_set_segment_kernel(int *seg, int n) { seg[0] = 0; seg[1] = n; }

static float af_limit_factor(float thr, float sigma) {
    // exp(-(d / dmax)^2 / (2 sigma^2)) > thr  <=>  (d / dmax)^2 < -2 sigma^2 ln(thr)
    if (thr >= 1.f) return -1.f;
    if (thr <= 0.f) return HUGE_VALF;
    return -2.f * sigma * sigma * logf(thr);
}

}  // namespace gcanet

using namespace gcanet;

extern "C" size_t gcanet_affinity_workspace_bytes(int segments) {
    return align_up((size_t)(2 * (segments > 0 ? segments : 1)) * sizeof(float));
}

// Largest squared pairwise distance per segment (rows lo..hi of x [n][C], seg [S + 1] int32 on the device) -> dmax2 [S].
extern "C" int gcanet_pairwise_max_distance(const float *x, const int32_t *seg, int n, int C, int segments, int max_segment,
                                            float *dmax2, gcanet_stream_t stream) {
    GCANET_REQUIRE(x && seg && dmax2 && n >= 1 && C >= 1 && segments >= 1 && segments <= 65535 && max_segment >= 1,
                   "pairwise_max_distance: bad arguments (n=%d C=%d segments=%d)", n, C, segments);
    cudaStream_t st = as_stream(stream);
    GCANET_CUDA_OK(cudaMemsetAsync(dmax2, 0, (size_t)segments * sizeof(float), st));
    const int tiles = ceil_div(max_segment, 64);
    GCANET_REQUIRE(tiles <= 65535, "pairwise_max_distance: segment of %d points is too large", max_segment);
    af_dmax_kernel<<<dim3(tiles, tiles, segments), 256, 0, st>>>(x, seg, C, reinterpret_cast<unsigned *>(dmax2));
    GCANET_LAUNCH_OK("af_dmax_kernel");
    return GCANET_OK;
}

// Replaces compute_batch_adjacency_matrix(x, dist_state=True, sigma) (M4:210-233) for one [n][C] cloud: adj [n][n].
extern "C" int gcanet_affinity_matrix(const float *x, int n, int C, float sigma, float *adj, void *ws, size_t ws_bytes,
                                      gcanet_stream_t stream) {
    GCANET_REQUIRE(x && adj && n >= 1 && C >= 1 && sigma > 0.f, "affinity_matrix: bad arguments (n=%d C=%d sigma=%g)", n, C, (double)sigma);
    GCANET_REQUIRE(ws && ws_bytes >= gcanet_affinity_workspace_bytes(1) + 16 && reinterpret_cast<uintptr_t>(ws) % kAlign == 0,
                   "affinity_matrix: workspace too small or misaligned");
    GCANET_REQUIRE(ceil_div(n, 64) <= 65535, "affinity_matrix: n=%d too large", n);
    cudaStream_t st = as_stream(stream);
    float *dmax2 = reinterpret_cast<float *>(ws);
    int *seg = reinterpret_cast<int *>(dmax2 + 2);
    af_set_segment_kernel<<<1, 1, 0, st>>>(seg, n);
    GCANET_LAUNCH_OK("af_set_segment_kernel");
    int rc = gcanet_pairwise_max_distance(x, seg, n, C, 1, n, dmax2, stream);
    if (rc) return rc;
    const int tiles = ceil_div(n, 64);
    af_matrix_kernel<<<dim3(tiles, tiles), 256, 0, st>>>(x, n, C, dmax2, 1.f / (2.f * sigma * sigma), adj);
    GCANET_LAUNCH_OK("af_matrix_kernel");
    return GCANET_OK;
}

static int af_run_query(AfQueryArgs &a, bool dense, long long *total, cudaStream_t st) {
    const unsigned grid = (unsigned)ceil_div64((long long)a.n * 32, 256);
    if (dense) af_query_kernel<0, true><<<grid, 256, 0, st>>>(a); else af_query_kernel<0, false><<<grid, 256, 0, st>>>(a);
    GCANET_LAUNCH_OK("af_query_kernel<count>");
    af_scan_kernel<<<1, 1024, 0, st>>>(a.start_len, a.n, total);
    GCANET_LAUNCH_OK("af_scan_kernel");
    if (dense) af_query_kernel<1, true><<<grid, 256, 0, st>>>(a); else af_query_kernel<1, false><<<grid, 256, 0, st>>>(a);
    GCANET_LAUNCH_OK("af_query_kernel<fill>");
    return GCANET_OK;
}

// Replaces ballquery_batch_p(coords, batch_idxs, batch_offsets, adj_mat_inst, thr_inst, adj_mat_para, thr_para, idx, start_len,
// n, meanActive, radius) (softgroup/ops/src/bfs_cluster/bfs_cluster.cpp:20-46, kernel bfs_cluster.cu:18-77) with the dense
// matrices the reference passes.  idx holds `capacity` ints; *total (device, int64) receives the number of neighbours found --
// when it exceeds capacity the caller re-runs with a larger buffer, as the reference's Python loop does (functions.py:460-472).
// Lists appear in point order (the reference: atomicAdd order), each in ascending neighbour index, at most 3000 per point.
extern "C" int gcanet_ball_query_dense(const float *xyz, const int32_t *batch_offsets, int n, int segments, const float *adj_inst,
                                       float thr_inst, const float *adj_para, float thr_para, float radius, int32_t *idx,
                                       long long capacity, int32_t *start_len, long long *total, gcanet_stream_t stream) {
    GCANET_REQUIRE(xyz && batch_offsets && adj_inst && adj_para && idx && start_len && total && n >= 1 && segments >= 1 && capacity >= 0,
                   "ball_query_dense: bad arguments");
    AfQueryArgs a{};
    a.xyz = xyz; a.seg = batch_offsets; a.adj_inst = adj_inst; a.adj_para = adj_para; a.thr_inst = thr_inst; a.thr_para = thr_para;
    a.r2 = radius * radius; a.n = n; a.S = segments; a.start_len = start_len; a.idx = idx; a.capacity = capacity;
    return af_run_query(a, true, total, as_stream(stream));
}

// The fused form of  compute_batch_adjacency_matrix(f_inst) + compute_batch_adjacency_matrix(f_para) + ball_query(...)
// (M4:1215-1233): no n x n matrix.  f_inst [n][Ci], f_para [n][Cp]; segments (clouds / classes) given by batch_offsets [S + 1],
// every segment normalised by its own largest pairwise distance, as one reference call per segment would.
extern "C" int gcanet_affinity_ball_query(const float *xyz, const int32_t *batch_offsets, int n, int segments, int max_segment,
                                          const float *f_inst, int Ci, float thr_inst, const float *f_para, int Cp, float thr_para,
                                          float sigma, float radius, int32_t *idx, long long capacity, int32_t *start_len,
                                          long long *total, void *ws, size_t ws_bytes, gcanet_stream_t stream) {
    GCANET_REQUIRE(xyz && batch_offsets && f_inst && f_para && idx && start_len && total && n >= 1 && segments >= 1 && Ci >= 1 && Cp >= 1 &&
                   sigma > 0.f && capacity >= 0, "affinity_ball_query: bad arguments");
    GCANET_REQUIRE(ws && ws_bytes >= gcanet_affinity_workspace_bytes(segments) && reinterpret_cast<uintptr_t>(ws) % kAlign == 0,
                   "affinity_ball_query: workspace too small or misaligned");
    float *dmax2 = reinterpret_cast<float *>(ws);
    AfQueryArgs a{};
    a.c2_inst = af_limit_factor(thr_inst, sigma);
    a.c2_para = af_limit_factor(thr_para, sigma);
    int rc = GCANET_OK;
    if (a.c2_inst >= 0.f && a.c2_inst != HUGE_VALF) rc = gcanet_pairwise_max_distance(f_inst, batch_offsets, n, Ci, segments, max_segment, dmax2, stream);
    if (rc) return rc;
    if (a.c2_para >= 0.f && a.c2_para != HUGE_VALF) rc = gcanet_pairwise_max_distance(f_para, batch_offsets, n, Cp, segments, max_segment, dmax2 + segments, stream);
    if (rc) return rc;
    a.xyz = xyz; a.seg = batch_offsets; a.f_inst = f_inst; a.f_para = f_para; a.Ci = Ci; a.Cp = Cp;
    a.dmax2_inst = dmax2; a.dmax2_para = dmax2 + segments; a.r2 = radius * radius; a.n = n; a.S = segments;
    a.start_len = start_len; a.idx = idx; a.capacity = capacity; a.include_self = (thr_inst < 0.f && thr_para < 0.f) ? 1 : 0;
    return af_run_query(a, false, total, as_stream(stream));
}
