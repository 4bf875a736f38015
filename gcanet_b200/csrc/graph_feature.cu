// Materialising edge-feature kernels (API-parity path) and the PN2 grouping operation.
//
// These exist so that get_graph_feature*/grouping_operation keep the reference's exact
// return shapes; the training fast path (edgeconv.cu) never forms these tensors.  They are
// pure data movement, bound by the HBM write of the [B][N][k][F] result: every output row
// is written with coalesced stores and neighbour rows are read as contiguous point-major
// rows from a staging copy that stays L2-resident (one cloud is N*C*4 bytes).
#include "common.cuh"

namespace gcanet {

// out[b][i][kk][0:C] = x_j - x_i, out[b][i][kk][C:2C] = x_i          (M4:120-123)
// C % 64 == 0: one warp per output row, four rows in flight; a lane moves 16 bytes per access (at C = 64 the row is exactly
// one 16-byte load pair and one 16-byte store per lane); rows < 2^31 (32-bit index arithmetic, checked by the caller)
__global__ void __launch_bounds__(256) edge_diff_center_kernel(const float *__restrict__ x_nc, const int64_t *__restrict__ idx,
                                                               float *__restrict__ out, int C, int N, int k, unsigned rows) {
    const int lane = threadIdx.x & 31;
    const unsigned nw = (gridDim.x * blockDim.x) >> 5;
    for (unsigned w0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w0 < rows; w0 += 4 * nw) {
        const float *xi[4], *xj[4];
        unsigned row[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            row[u] = min(w0 + u * nw, rows - 1);                     // (clamped rows are recomputed, not written)
            const unsigned bi = row[u] / (unsigned)k, b = bi / (unsigned)N;
            xi[u] = x_nc + (size_t)bi * C;
            xj[u] = x_nc + ((size_t)b * N + (size_t)idx[row[u]]) * C;
        }
        for (int o4 = lane * 4; o4 < 2 * C; o4 += 128) {              // o4: offset inside the 2C-float output row
            const bool diff = o4 < C;
            const int c4 = diff ? o4 : o4 - C;
            float4 ci[4], cj[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                ci[u] = __ldg(reinterpret_cast<const float4 *>(xi[u] + c4));
                if (diff) cj[u] = __ldg(reinterpret_cast<const float4 *>(xj[u] + c4));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (w0 + u * nw >= rows) continue;
                float4 v = ci[u];
                if (diff) v = make_float4(cj[u].x - ci[u].x, cj[u].y - ci[u].y, cj[u].z - ci[u].z, cj[u].w - ci[u].w);
                *reinterpret_cast<float4 *>(out + (size_t)row[u] * 2 * C + o4) = v;
            }
        }
    }
}

// any other C (the xyz layers, C = 3 / 6): one thread per output element, so a warp writes 128 contiguous bytes spanning
// several rows
template <typename I>
__global__ void __launch_bounds__(256) edge_diff_center_small_kernel(const float *__restrict__ x_nc, const int64_t *__restrict__ idx,
                                                                     float *__restrict__ out, int C, int N, int k, I total) {
    const I w2 = 2 * C, nt = (I)gridDim.x * blockDim.x;
    for (I t = (I)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += nt) {
        const I row = t / w2;
        const int c2 = (int)(t - row * w2);
        const I bi = row / (I)k;
        const float ci = x_nc[(size_t)bi * C + (c2 < C ? c2 : c2 - C)];
        float v = ci;
        if (c2 < C) {
            const I b = bi / (I)N;
            v = x_nc[((size_t)b * N + (size_t)idx[row]) * C + c2] - ci;
        }
        out[t] = v;
    }
}

// out[b][i][kk] = (clamp(n_i.n_j, -.99, .99), n_j - n_i, n_i), x has 6 channels   (M4:189-204)
__global__ void edge_normal_angle_kernel(const float *__restrict__ x_nc, const int64_t *__restrict__ idx,
                                         float *__restrict__ out, int N, int k, long long rows) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nt = (long long)gridDim.x * blockDim.x;
    for (; t < rows; t += nt) {
        long long bi = t / k;
        long long b = bi / N;
        long long j = idx[t];
        const float *ni = x_nc + bi * 6 + 3;
        const float *nj = x_nc + (b * N + j) * 6 + 3;
        float a0 = ni[0], a1 = ni[1], a2 = ni[2];
        float b0 = nj[0], b1 = nj[1], b2 = nj[2];
        // reference: elementwise product then sum over the 3 channels (no FMA), then clamp
        float dot = __fadd_rn(__fadd_rn(__fmul_rn(a0, b0), __fmul_rn(a1, b1)), __fmul_rn(a2, b2));
        dot = fminf(fmaxf(dot, -0.99f), 0.99f);
        float *o = out + t * 7;
        o[0] = dot;
        o[1] = b0 - a0; o[2] = b1 - a1; o[3] = b2 - a2;
        o[4] = a0; o[5] = a1; o[6] = a2;
    }
}

// grad of edge_diff_center w.r.t. x, accumulated point-major into g_nc (zeroed by the caller)
__global__ void edge_diff_center_grad_kernel(const float *__restrict__ go, const int64_t *__restrict__ idx,
                                             float *__restrict__ g_nc, int C, int N, int k, long long points) {
    // one warp per (b, i)
    const int lane = threadIdx.x & 31;
    long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (; w < points; w += nw) {
        long long b = w / N;
        for (int c0 = 0; c0 < C; c0 += 32) {
            int c = c0 + lane;
            float self = 0.f;
            for (int kk = 0; kk < k; ++kk) {
                long long row = w * k + kk;
                long long j = idx[row];
                if (c < C) {
                    float gd = go[row * 2 * C + c];
                    self += go[row * 2 * C + C + c] - gd;
                    atomicAdd(g_nc + (b * N + j) * C + c, gd);
                }
            }
            if (c < C) atomicAdd(g_nc + w * C + c, self);
        }
    }
}

// grad of edge_normal_angle w.r.t. x (channels 3..5 only; xyz channels get zero)
__global__ void edge_normal_angle_grad_kernel(const float *__restrict__ go, const float *__restrict__ x_nc,
                                              const int64_t *__restrict__ idx, float *__restrict__ g_nc,
                                              int N, int k, long long rows) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nt = (long long)gridDim.x * blockDim.x;
    for (; t < rows; t += nt) {
        long long bi = t / k;
        long long b = bi / N;
        long long j = idx[t];
        const float *ni = x_nc + bi * 6 + 3;
        const float *nj = x_nc + (b * N + j) * 6 + 3;
        float a0 = ni[0], a1 = ni[1], a2 = ni[2];
        float b0 = nj[0], b1 = nj[1], b2 = nj[2];
        float dot = __fadd_rn(__fadd_rn(__fmul_rn(a0, b0), __fmul_rn(a1, b1)), __fmul_rn(a2, b2));
        const float *g = go + t * 7;
        float ga = (dot >= -0.99f && dot <= 0.99f) ? g[0] : 0.f;   // clamp passes the gradient inside [min,max]
        float *gi = g_nc + bi * 6 + 3;
        float *gj = g_nc + (b * N + j) * 6 + 3;
        atomicAdd(gi + 0, ga * b0 - g[1] + g[4]);
        atomicAdd(gi + 1, ga * b1 - g[2] + g[5]);
        atomicAdd(gi + 2, ga * b2 - g[3] + g[6]);
        atomicAdd(gj + 0, ga * a0 + g[1]);
        atomicAdd(gj + 1, ga * a1 + g[2]);
        atomicAdd(gj + 2, ga * a2 + g[3]);
    }
}

// out[b][c][j][s] = points[b][c][idx[b][j][s]]          (group_points_gpu.cu:20-27)
// One thread per V consecutive (j, s) slots of a batch element: it reads its V indices once and walks the channels, so
// the index array is read once instead of once per channel, the stores are V-wide and coalesced across the warp, and the
// inner loop has no division.  (The reference launches one CTA per batch element.)
template <int V>
__global__ void __launch_bounds__(256) group_points_kernel(const float *__restrict__ points, const int32_t *__restrict__ idx,
                                                           float *__restrict__ out, int c, int n, long long ms, long long work) {
    const long long per_b = ms / V;
    const long long nt = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < work; t += nt) {
        const long long b = t / per_b, e = (t - b * per_b) * V;
        int ii[V];
#pragma unroll
        for (int v = 0; v < V; ++v) ii[v] = idx[b * ms + e + v];
        const float *src = points + b * c * n;
        float *dst = out + b * c * ms + e;
#pragma unroll 4
        for (int l = 0; l < c; ++l) {
            if constexpr (V == 4) {
                *reinterpret_cast<float4 *>(dst + (long long)l * ms) =
                    make_float4(__ldg(src + (long long)l * n + ii[0]), __ldg(src + (long long)l * n + ii[1]),
                                __ldg(src + (long long)l * n + ii[2]), __ldg(src + (long long)l * n + ii[3]));
            } else {
                dst[(long long)l * ms] = __ldg(src + (long long)l * n + ii[0]);
            }
        }
    }
}

// grad_points[b][c][idx[b][j][s]] += grad_out[b][c][j][s]   (group_points_gpu.cu:56-62)
__global__ void __launch_bounds__(256) group_points_grad_kernel(const float *__restrict__ go, const int32_t *__restrict__ idx,
                                                                float *__restrict__ gp, int c, int n, long long ms, long long work) {
    const long long nt = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < work; t += nt) {
        const long long b = t / ms, e = t - b * ms;
        const int ii = idx[t];
        const float *src = go + b * c * ms + e;
        float *dst = gp + b * c * n + ii;
#pragma unroll 4
        for (int l = 0; l < c; ++l) atomicAdd(dst + (long long)l * n, __ldg(src + (long long)l * ms));
    }
}

static int grid_for(long long work_items, int per_block) {
    long long g = (work_items + per_block - 1) / per_block;
    long long cap = (long long)kNumSMs * 32;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace gcanet

using namespace gcanet;

extern "C" int gcanet_graph_feature_channels(int C, int variant) {
    if (variant == GCANET_EDGE_DIFF_CENTER) return 2 * C;
    if (variant == GCANET_EDGE_NORMAL_ANGLE) return 7;
    return 0;
}

extern "C" size_t gcanet_graph_feature_workspace_bytes(int B, int C, int N, int k, int variant) {
    (void)k; (void)variant;
    return align_up((size_t)B * N * C * sizeof(float));
}

extern "C" size_t gcanet_graph_feature_grad_workspace_bytes(int B, int C, int N, int k, int variant) {
    (void)k; (void)variant;
    return 2 * align_up((size_t)B * N * C * sizeof(float));
}

static int check_gf(const char *who, int B, int C, int N, int k, int variant) {
    GCANET_REQUIRE(B >= 1 && C >= 1 && N >= 1 && k >= 1, "%s: bad shape B=%d C=%d N=%d k=%d", who, B, C, N, k);
    GCANET_REQUIRE(variant == GCANET_EDGE_DIFF_CENTER || variant == GCANET_EDGE_NORMAL_ANGLE, "%s: bad variant %d", who, variant);
    GCANET_REQUIRE(variant != GCANET_EDGE_NORMAL_ANGLE || C == 6, "%s: the normal-angle feature needs C = 6 (got %d)", who, C);
    return GCANET_OK;
}

extern "C" int gcanet_graph_feature(const float *x, const int64_t *idx, float *out, int B, int C, int N, int k,
                                    int variant, void *ws, size_t ws_bytes, gcanet_stream_t stream) {
    GCANET_REQUIRE(x && idx && out, "graph_feature: null pointer");
    int rc = check_gf("graph_feature", B, C, N, k, variant);
    if (rc) return rc;
    if (!ws || ws_bytes < gcanet_graph_feature_workspace_bytes(B, C, N, k, variant) ||
        reinterpret_cast<uintptr_t>(ws) % kAlign) {
        set_error("graph_feature: workspace too small or misaligned");
        return GCANET_ERR_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    float *x_nc = static_cast<float *>(ws);
    rc = launch_cn_to_nc(x, x_nc, B, C, N, C, st);
    if (rc) return rc;
    long long rows = (long long)B * N * k;
    if (variant == GCANET_EDGE_DIFF_CENTER) {
        const long long total = rows * 2 * C;
        if (C % 64 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0 && rows < 2147483647ll)
            edge_diff_center_kernel<<<grid_for(rows, 32), 256, 0, st>>>(x_nc, idx, out, C, N, k, (unsigned)rows);
        else if (total < 4294967295ll)
            edge_diff_center_small_kernel<unsigned><<<grid_for(total, 256), 256, 0, st>>>(x_nc, idx, out, C, N, k, (unsigned)total);
        else
            edge_diff_center_small_kernel<long long><<<grid_for(total, 256), 256, 0, st>>>(x_nc, idx, out, C, N, k, total);
        GCANET_LAUNCH_OK("edge_diff_center_kernel");
    } else {
        edge_normal_angle_kernel<<<grid_for(rows, 256), 256, 0, st>>>(x_nc, idx, out, N, k, rows);
        GCANET_LAUNCH_OK("edge_normal_angle_kernel");
    }
    return GCANET_OK;
}

extern "C" int gcanet_graph_feature_grad(const float *grad_out, const float *x, const int64_t *idx, float *grad_x,
                                         int B, int C, int N, int k, int variant, void *ws, size_t ws_bytes,
                                         gcanet_stream_t stream) {
    GCANET_REQUIRE(grad_out && idx && grad_x, "graph_feature_grad: null pointer");
    int rc = check_gf("graph_feature_grad", B, C, N, k, variant);
    if (rc) return rc;
    GCANET_REQUIRE(variant != GCANET_EDGE_NORMAL_ANGLE || x != nullptr, "graph_feature_grad: x is needed for the normal-angle variant");
    if (!ws || ws_bytes < gcanet_graph_feature_grad_workspace_bytes(B, C, N, k, variant) ||
        reinterpret_cast<uintptr_t>(ws) % kAlign) {
        set_error("graph_feature_grad: workspace too small or misaligned");
        return GCANET_ERR_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    Carver cv(ws);
    float *g_nc = cv.take<float>((size_t)B * N * C);
    float *x_nc = cv.take<float>((size_t)B * N * C);
    GCANET_CUDA_OK(cudaMemsetAsync(g_nc, 0, (size_t)B * N * C * sizeof(float), st));
    if (variant == GCANET_EDGE_DIFF_CENTER) {
        long long pts = (long long)B * N;
        edge_diff_center_grad_kernel<<<grid_for(pts, 8), 256, 0, st>>>(grad_out, idx, g_nc, C, N, k, pts);
        GCANET_LAUNCH_OK("edge_diff_center_grad_kernel");
    } else {
        rc = launch_cn_to_nc(x, x_nc, B, C, N, C, st);
        if (rc) return rc;
        long long rows = (long long)B * N * k;
        edge_normal_angle_grad_kernel<<<grid_for(rows, 256), 256, 0, st>>>(grad_out, x_nc, idx, g_nc, N, k, rows);
        GCANET_LAUNCH_OK("edge_normal_angle_grad_kernel");
    }
    return launch_nc_to_cn(g_nc, grad_x, B, C, N, C, st);
}

extern "C" int gcanet_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                                   const int32_t *idx, float *out, gcanet_stream_t stream) {
    GCANET_REQUIRE(points && idx && out, "group_points: null pointer");
    GCANET_REQUIRE(b >= 1 && c >= 1 && n >= 1 && npoints >= 1 && nsample >= 1, "group_points: bad shape");
    const long long ms = (long long)npoints * nsample;
    if (ms % 4 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0) {
        const long long work = (long long)b * (ms / 4);
        group_points_kernel<4><<<grid_for(work, 256), 256, 0, as_stream(stream)>>>(points, idx, out, c, n, ms, work);
    } else {
        const long long work = (long long)b * ms;
        group_points_kernel<1><<<grid_for(work, 256), 256, 0, as_stream(stream)>>>(points, idx, out, c, n, ms, work);
    }
    GCANET_LAUNCH_OK("group_points_kernel");
    return GCANET_OK;
}

extern "C" int gcanet_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                                        const int32_t *idx, float *grad_points, gcanet_stream_t stream) {
    GCANET_REQUIRE(grad_out && idx && grad_points, "group_points_grad: null pointer");
    GCANET_REQUIRE(b >= 1 && c >= 1 && n >= 1 && npoints >= 1 && nsample >= 1, "group_points_grad: bad shape");
    cudaStream_t st = as_stream(stream);
    GCANET_CUDA_OK(cudaMemsetAsync(grad_points, 0, (size_t)b * c * n * sizeof(float), st));
    const long long ms = (long long)npoints * nsample, work = (long long)b * ms;
    group_points_grad_kernel<<<grid_for(work, 256), 256, 0, st>>>(grad_out, idx, grad_points, c, n, ms, work);
    GCANET_LAUNCH_OK("group_points_grad_kernel");
    return GCANET_OK;
}
