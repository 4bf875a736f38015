// GroupNorm + ReLU of the per-point heads, channel-major [B][C][N] (what the 1x1 convolutions produce):
//     y = act(GroupNorm(x))          F.relu(self.bn1(self.conv1(x))) and its siblings, M4:644-713
// forward and backward.  All of it is HBM streaming; the point of the kernels is the grid: a group's (C / G) * N values
// are contiguous in this layout, so its statistics are split over many CTAs (fp64 partials, fixed-order finalize) instead
// of one CTA per (sample, group) row, which leaves a B200 with B * G = 64..128 busy CTAs for 2.6 MB rows.
//
//   forward   gn_moments_kernel (read x once) -> gn_stats_finalize_kernel -> gn_apply_kernel (read x, write y)
//   backward  gn_bwd_rows_kernel   per (b, c): r1 = sum_n dz, r2 = sum_n dz xhat            (read dy, x)
//             gn_bwd_combine_kernel             dgamma, dbeta, per (b, g): s1 = sum gamma r1, s2 = sum gamma r2
//             gn_bwd_apply_kernel  dx = rstd (gamma dz - s1 / M - xhat s2 / M)                (read dy, x, write dx)
//   with dz = dy * act'(y), y recomputed from x (nothing but the statistics is saved).
#include "common.cuh"

namespace gcanet {

constexpr int GN_THREADS = 256;
constexpr int GN_CHUNK = 16384;        // values per CTA of the moments pass

struct GnArgs {
    const float *x;        // [B][C][N]
    const float *gamma, *beta;
    float *y;
    float *stats;          // [B][G][2] mean, rstd
    double *part;          // [B][G][chunks][2]
    int B, C, N, G, act, chunks;
    float eps;
};

__device__ __forceinline__ double gn_block_sum(double v, double *red) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < GN_THREADS / 32; ++w) s += red[w];
    return s;
}

// grid (chunks, B * G)
__global__ void __launch_bounds__(GN_THREADS) gn_moments_kernel(GnArgs a) {
    __shared__ double red[GN_THREADS / 32];
    const size_t M = (size_t)(a.C / a.G) * a.N;
    const float *row = a.x + (size_t)blockIdx.y * M;
    const size_t lo = (size_t)blockIdx.x * GN_CHUNK, hi = min(lo + GN_CHUNK, M);
    float s1 = 0.f, s2 = 0.f;                       // <= 64 values per thread: fp32 is exact enough, fp64 from there on
    if ((M & 3) == 0 && (reinterpret_cast<uintptr_t>(row) & 15) == 0) {
        for (size_t e = lo + threadIdx.x * 4; e < hi; e += GN_THREADS * 4) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(row + e));
            s1 += (v.x + v.y) + (v.z + v.w);
            s2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s2))));
        }
    } else {
        for (size_t e = lo + threadIdx.x; e < hi; e += GN_THREADS) { const float v = row[e]; s1 += v; s2 = fmaf(v, v, s2); }
    }
    const double t1 = gn_block_sum((double)s1, red), t2 = gn_block_sum((double)s2, red);
    if (threadIdx.x == 0) {
        double *p = a.part + ((size_t)blockIdx.y * a.chunks + blockIdx.x) * 2;
        p[0] = t1; p[1] = t2;
    }
}

// one warp per (b, g)
__global__ void gn_stats_finalize_kernel(GnArgs a) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= a.B * a.G) return;
    double s1 = 0.0, s2 = 0.0;
    for (int c = lane; c < a.chunks; c += 32) {
        s1 += a.part[((size_t)row * a.chunks + c) * 2];
        s2 += a.part[((size_t)row * a.chunks + c) * 2 + 1];
    }
    for (int o = 16; o; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    if (lane == 0) {
        const double M = (double)(a.C / a.G) * a.N;
        const double mean = s1 / M;
        double var = s2 / M - mean * mean;             // biased, like torch
        if (var < 0.0) var = 0.0;
        a.stats[row * 2] = (float)mean;
        a.stats[row * 2 + 1] = (float)(1.0 / sqrt(var + (double)a.eps));
    }
}

// grid (B * C, ceil(N / (8 * GN_THREADS)))
__global__ void __launch_bounds__(GN_THREADS) gn_apply_kernel(GnArgs a) {
    const int bc = blockIdx.x, b = bc / a.C, c = bc % a.C, g = c / (a.C / a.G);
    const float mean = a.stats[(b * a.G + g) * 2], rstd = a.stats[(b * a.G + g) * 2 + 1];
    const float sc = rstd * a.gamma[c], sh = fmaf(-mean, sc, a.beta[c]);
    const float *x = a.x + (size_t)bc * a.N;
    float *y = a.y + (size_t)bc * a.N;
    const bool relu = a.act == 1;
    if ((a.N & 3) == 0) {
        for (int e = (blockIdx.y * GN_THREADS + threadIdx.x) * 4; e < a.N; e += gridDim.y * GN_THREADS * 4) {
            float4 v = __ldg(reinterpret_cast<const float4 *>(x + e));
            v.x = fmaf(v.x, sc, sh); v.y = fmaf(v.y, sc, sh); v.z = fmaf(v.z, sc, sh); v.w = fmaf(v.w, sc, sh);
            if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            *reinterpret_cast<float4 *>(y + e) = v;
        }
    } else {
        for (int e = blockIdx.y * GN_THREADS + threadIdx.x; e < a.N; e += gridDim.y * GN_THREADS) {
            const float v = fmaf(x[e], sc, sh);
            y[e] = relu ? fmaxf(v, 0.f) : v;
        }
    }
}

struct GnBwdArgs {
    GnArgs f;
    const float *dy;       // [B][C][N]
    float *dx;             // [B][C][N]
    float *rows;           // [B][C][2]   r1 = sum_n dz, r2 = sum_n dz xhat
    float *sg;             // [B][G][2]   s1, s2 (already divided by M)
    float *dgamma, *dbeta; // [C]
};

// grid (B * C): one CTA per (b, c) row
__global__ void __launch_bounds__(GN_THREADS) gn_bwd_rows_kernel(GnBwdArgs p) {
    __shared__ double red[GN_THREADS / 32];
    const GnArgs &a = p.f;
    const int bc = blockIdx.x, b = bc / a.C, c = bc % a.C, g = c / (a.C / a.G);
    const float mean = a.stats[(b * a.G + g) * 2], rstd = a.stats[(b * a.G + g) * 2 + 1];
    const float sc = rstd * a.gamma[c], sh = fmaf(-mean, sc, a.beta[c]);
    const float *x = a.x + (size_t)bc * a.N, *dy = p.dy + (size_t)bc * a.N;
    const bool relu = a.act == 1;
    float r1 = 0.f, r2 = 0.f;
    auto one = [&](float xv, float dv) {
        const float dz = (relu && !(fmaf(xv, sc, sh) > 0.f)) ? 0.f : dv;
        r1 += dz;
        r2 = fmaf(dz, (xv - mean) * rstd, r2);
    };
    if ((a.N & 3) == 0) {
        for (int e = threadIdx.x * 4; e < a.N; e += GN_THREADS * 4) {
            const float4 xv = __ldg(reinterpret_cast<const float4 *>(x + e)), dv = __ldg(reinterpret_cast<const float4 *>(dy + e));
            one(xv.x, dv.x); one(xv.y, dv.y); one(xv.z, dv.z); one(xv.w, dv.w);
        }
    } else {
        for (int e = threadIdx.x; e < a.N; e += GN_THREADS) one(x[e], dy[e]);
    }
    const double t1 = gn_block_sum((double)r1, red), t2 = gn_block_sum((double)r2, red);
    if (threadIdx.x == 0) { p.rows[bc * 2] = (float)t1; p.rows[bc * 2 + 1] = (float)t2; }
}

// grid 1 + B * G blocks of 128 threads: block 0 .. : channels (dgamma, dbeta); the rest: one (b, g) each
__global__ void __launch_bounds__(128) gn_bwd_combine_kernel(GnBwdArgs p, int chan_blocks) {
    const GnArgs &a = p.f;
    if ((int)blockIdx.x < chan_blocks) {
        const int c = blockIdx.x * 128 + threadIdx.x;
        if (c >= a.C) return;
        double g1 = 0.0, g2 = 0.0;
        for (int b = 0; b < a.B; ++b) { g1 += p.rows[(b * a.C + c) * 2]; g2 += p.rows[(b * a.C + c) * 2 + 1]; }
        p.dbeta[c] = (float)g1;
        p.dgamma[c] = (float)g2;
        return;
    }
    __shared__ double red[2][4];
    const int bg = blockIdx.x - chan_blocks, b = bg / a.G, g = bg % a.G, cpg = a.C / a.G;
    double s1 = 0.0, s2 = 0.0;
    for (int i = threadIdx.x; i < cpg; i += 128) {
        const int c = g * cpg + i;
        const double gm = a.gamma[c];
        s1 += gm * p.rows[(b * a.C + c) * 2];
        s2 += gm * p.rows[(b * a.C + c) * 2 + 1];
    }
    for (int o = 16; o; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double M = (double)cpg * a.N;
        p.sg[bg * 2] = (float)((red[0][0] + red[0][1] + red[0][2] + red[0][3]) / M);
        p.sg[bg * 2 + 1] = (float)((red[1][0] + red[1][1] + red[1][2] + red[1][3]) / M);
    }
}

// grid (B * C, ceil(N / (8 * GN_THREADS)))
__global__ void __launch_bounds__(GN_THREADS) gn_bwd_apply_kernel(GnBwdArgs p) {
    const GnArgs &a = p.f;
    const int bc = blockIdx.x, b = bc / a.C, c = bc % a.C, g = c / (a.C / a.G);
    const float mean = a.stats[(b * a.G + g) * 2], rstd = a.stats[(b * a.G + g) * 2 + 1];
    const float gm = a.gamma[c];
    const float sc = rstd * gm, sh = fmaf(-mean, sc, a.beta[c]);
    const float s1 = p.sg[(b * a.G + g) * 2], s2 = p.sg[(b * a.G + g) * 2 + 1];
    const float *x = a.x + (size_t)bc * a.N, *dy = p.dy + (size_t)bc * a.N;
    float *dx = p.dx + (size_t)bc * a.N;
    const bool relu = a.act == 1;
    auto one = [&](float xv, float dv) {
        const float dz = (relu && !(fmaf(xv, sc, sh) > 0.f)) ? 0.f : dv;
        return rstd * (fmaf(gm, dz, -s1) - (xv - mean) * rstd * s2);
    };
    if ((a.N & 3) == 0) {
        for (int e = (blockIdx.y * GN_THREADS + threadIdx.x) * 4; e < a.N; e += gridDim.y * GN_THREADS * 4) {
            const float4 xv = __ldg(reinterpret_cast<const float4 *>(x + e)), dv = __ldg(reinterpret_cast<const float4 *>(dy + e));
            *reinterpret_cast<float4 *>(dx + e) = make_float4(one(xv.x, dv.x), one(xv.y, dv.y), one(xv.z, dv.z), one(xv.w, dv.w));
        }
    } else {
        for (int e = blockIdx.y * GN_THREADS + threadIdx.x; e < a.N; e += gridDim.y * GN_THREADS) dx[e] = one(x[e], dy[e]);
    }
}

static int gn_check(const gcanet_group_norm_desc *d) {
    GCANET_REQUIRE(d != nullptr, "group_norm: null descriptor");
    GCANET_REQUIRE(d->B >= 1 && d->C >= 1 && d->N >= 1 && (long long)d->B * d->C <= 2147483647ll / 2, "group_norm: bad shape B=%d C=%d N=%d",
                   d->B, d->C, d->N);
    GCANET_REQUIRE(d->groups >= 1 && d->C % d->groups == 0, "group_norm: groups=%d must divide C=%d", d->groups, d->C);
    GCANET_REQUIRE((long long)d->B * d->groups <= 65535, "group_norm: B * groups = %lld exceeds 65535", (long long)d->B * d->groups);
    GCANET_REQUIRE(d->eps > 0.f && (d->act == 0 || d->act == 1), "group_norm: eps must be positive, act 0 (none) or 1 (ReLU)");
    return GCANET_OK;
}

static int gn_chunks(const gcanet_group_norm_desc *d) {
    const size_t M = (size_t)(d->C / d->groups) * d->N;
    return (int)((M + GN_CHUNK - 1) / GN_CHUNK);
}

struct GnWs { double *part; float *rows, *sg; };
static size_t gn_plan_ws(const gcanet_group_norm_desc *d, void *base, GnWs *w) {
    Carver cv(base);
    GnWs t;
    t.part = cv.take<double>((size_t)d->B * d->groups * gn_chunks(d) * 2);
    t.rows = cv.take<float>((size_t)d->B * d->C * 2);
    t.sg = cv.take<float>((size_t)d->B * d->groups * 2);
    if (w) *w = t;
    return cv.off;
}

static GnArgs gn_args(const gcanet_group_norm_desc *d, const float *x, const float *gamma, const float *beta, float *y, float *stats,
                      double *part) {
    GnArgs a{};
    a.x = x; a.gamma = gamma; a.beta = beta; a.y = y; a.stats = stats; a.part = part;
    a.B = d->B; a.C = d->C; a.N = d->N; a.G = d->groups; a.act = d->act; a.chunks = gn_chunks(d); a.eps = d->eps;
    return a;
}

}  // namespace gcanet

using namespace gcanet;

extern "C" size_t gcanet_group_norm_workspace_bytes(const gcanet_group_norm_desc *d) {
    if (d == nullptr || d->B < 1 || d->C < 1 || d->N < 1 || d->groups < 1 || d->C % d->groups) return 0;
    return gn_plan_ws(d, nullptr, nullptr);
}

extern "C" int gcanet_group_norm_forward(const gcanet_group_norm_desc *d, const float *x, const float *gamma, const float *beta,
                                         float *y, float *stats, void *ws, size_t ws_bytes, gcanet_stream_t stream) {
    int rc = gn_check(d);
    if (rc) return rc;
    GCANET_REQUIRE(x && gamma && beta && y && stats, "group_norm_forward: null pointer");
    if (ws == nullptr || ws_bytes < gn_plan_ws(d, nullptr, nullptr) || reinterpret_cast<uintptr_t>(ws) % kAlign) {
        set_error("group_norm_forward: workspace too small or misaligned (%zu given)", ws_bytes);
        return GCANET_ERR_WORKSPACE;
    }
    GnWs w;
    gn_plan_ws(d, ws, &w);
    cudaStream_t st = as_stream(stream);
    const GnArgs a = gn_args(d, x, gamma, beta, y, stats, w.part);
    gn_moments_kernel<<<dim3(a.chunks, d->B * d->groups), GN_THREADS, 0, st>>>(a);
    GCANET_LAUNCH_OK("gn_moments_kernel");
    gn_stats_finalize_kernel<<<ceil_div(d->B * d->groups, 4), 128, 0, st>>>(a);
    GCANET_LAUNCH_OK("gn_stats_finalize_kernel");
    gn_apply_kernel<<<dim3(d->B * d->C, min(ceil_div(d->N, GN_THREADS * 8), 65535)), GN_THREADS, 0, st>>>(a);
    GCANET_LAUNCH_OK("gn_apply_kernel");
    return GCANET_OK;
}

extern "C" int gcanet_group_norm_backward(const gcanet_group_norm_desc *d, const float *x, const float *gamma, const float *beta,
                                          const float *stats, const float *grad_y, float *grad_x, float *grad_gamma,
                                          float *grad_beta, void *ws, size_t ws_bytes, gcanet_stream_t stream) {
    int rc = gn_check(d);
    if (rc) return rc;
    GCANET_REQUIRE(x && gamma && beta && stats && grad_y && grad_x && grad_gamma && grad_beta, "group_norm_backward: null pointer");
    if (ws == nullptr || ws_bytes < gn_plan_ws(d, nullptr, nullptr) || reinterpret_cast<uintptr_t>(ws) % kAlign) {
        set_error("group_norm_backward: workspace too small or misaligned (%zu given)", ws_bytes);
        return GCANET_ERR_WORKSPACE;
    }
    GnWs w;
    gn_plan_ws(d, ws, &w);
    cudaStream_t st = as_stream(stream);
    GnBwdArgs p{};
    p.f = gn_args(d, x, gamma, beta, nullptr, const_cast<float *>(stats), w.part);
    p.dy = grad_y; p.dx = grad_x; p.rows = w.rows; p.sg = w.sg; p.dgamma = grad_gamma; p.dbeta = grad_beta;
    gn_bwd_rows_kernel<<<d->B * d->C, GN_THREADS, 0, st>>>(p);
    GCANET_LAUNCH_OK("gn_bwd_rows_kernel");
    const int chan_blocks = ceil_div(d->C, 128);
    gn_bwd_combine_kernel<<<chan_blocks + d->B * d->groups, 128, 0, st>>>(p, chan_blocks);
    GCANET_LAUNCH_OK("gn_bwd_combine_kernel");
    gn_bwd_apply_kernel<<<dim3(d->B * d->C, min(ceil_div(d->N, GN_THREADS * 8), 65535)), GN_THREADS, 0, st>>>(p);
    GCANET_LAUNCH_OK("gn_bwd_apply_kernel");
    return GCANET_OK;
}
