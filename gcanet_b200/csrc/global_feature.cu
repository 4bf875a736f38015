// Encoder tail of the DGCNN backbone (M4:507-511): the consumer of x1 | x2 | x3.
//
//   x = relu(GroupNorm(G, Cout)(Conv1d(K -> Cout, 1)(x_features)));  x4 = max over the N points        (K = 256, Cout = 1024)
//
// The reference materialises the [B][Cout][N] activation three times (conv, norm, relu: 655 MB each at B = 16, N = 10^4)
// and then repeats x4 N times into a [B][1280][N] tensor.  Here the activation never exists:
//
//   forward   one tcgen05 GEMM  D[c][n] = W[c][:] . x_n  with the output channels on the TMEM lanes, so that an epilogue
//             thread owns ONE channel and reduces over the points of its tile in registers: running max of sign(gamma) y
//             (+ the point that attains it), sum y, sum y^2.  GroupNorm is affine per (cloud, channel) with the sign of
//             gamma and ReLU is non-decreasing, hence  max_n relu(GN(y_cn)) = relu(GN(gamma >= 0 ? max_n y : min_n y)).
//             W (bf16 hi | lo split) lives in TMEM as the A operand (TS-mode MMA), the points are split into hi | lo
//             while they are staged into shared memory (as in gemm_tc.cu): hi*hi + hi*lo + lo*hi, fp32 accumulate.
//   backward  dy_cn = [n = n*_c] s_c + A_g + K_g y_cn  (A_g, K_g: GroupNorm's mean / variance paths), and because
//             y_cn = W_c . x_n + b_c is linear in x_n the dense part collapses onto per-cloud K x K matrices:
//               dX_n = M_b x_n + c_b + sum_{c : n*_c = n} s_c W_c,     M_b = sum_g K_bg W_g^T W_g
//               dW_c = sum_b [ s_bc x_{n*} + (A_bg + K_bg b_c) Sx_b + K_bg W_c G_b ],   G_b = X_b^T X_b,  Sx_b = sum_n x_n
//             i.e. two batched tensor-core GEMMs of K x K weights (gemm_tc.cu) and a handful of small kernels; nothing of
//             size Cout x N is ever formed.
#include "common.cuh"
#include "tc_ptx.cuh"

#include <math_constants.h>

namespace gcanet {

int gemm_tc_batched_try(const float *A, int lda, long long batch_a, const float *Bt, int ldb, long long batch_b, float *C, int ldc,
                        long long batch_c, const float *cbias, int batch_bias, int M, int N, int K, int batches, cudaStream_t st);
int gemm_tn_tc_batched(const float *X, int ldx, long long batch_x, const float *Y, int ldy, long long batch_y, float *part, int M,
                       int splits, int batches, cudaStream_t st);

constexpr int GF_K = 256;             // input channels (64 + 64 + 128)
constexpr int GF_BM = 128;            // output channels per CTA (UMMA M, TMEM lanes)
constexpr int GF_BN = 128;            // points per tile (UMMA N)
constexpr int GF_KB = 64;             // K chunk per pipeline stage (one 128-byte swizzle row of bf16)
constexpr int GF_KCH = GF_K / GF_KB;
constexpr int GF_STAGES = 5;          // 32 KB each: hi | lo tile of 128 points x 64 channels
constexpr int GF_LOADERS = 256;
constexpr int GF_THREADS = GF_LOADERS + 128 + 32;
constexpr int GF_TILE = GF_BN * 128;  // one [128][64] bf16 tile
constexpr int GF_STAGE = 2 * GF_TILE;
constexpr int GF_GRAM_SPLITS = 8;

__device__ __forceinline__ uint32_t gf_sw128(int r, int kc) { return (uint32_t)(r * 128 + ((((kc >> 3) ^ (r & 7)) << 4) | ((kc & 7) << 1))); }

__device__ __forceinline__ void gf_split4(float4 v, uint2 &hw, uint2 &lw) {
    const float f[4] = {v.x, v.y, v.z, v.w};
    __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        h[i] = __float2bfloat16_rn(f[i]);
        l[i] = __float2bfloat16_rn(f[i] - __bfloat162float(h[i]));
    }
    hw.x = (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16);
    hw.y = (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16);
    lw.x = (uint32_t)__bfloat16_as_ushort(l[0]) | ((uint32_t)__bfloat16_as_ushort(l[1]) << 16);
    lw.y = (uint32_t)__bfloat16_as_ushort(l[2]) | ((uint32_t)__bfloat16_as_ushort(l[3]) << 16);
}

struct GfFwdArgs {
    const float *x;        // [B][N][K] point-major
    const float *w;        // [Cout][K]
    const float *gamma;    // [Cout] only the sign is used
    float4 *part;          // [B][S][Cout] (zmax, arg as int bits, sum, sum of squares) of acc = W_c . x_n (no bias)
    int N, Cout, S, tiles;
};

// grid (S, Cout / 128, B): CTA = 128 output channels x the point tiles s, s + S, ... of one cloud
__global__ void __launch_bounds__(GF_THREADS, 1) gf_forward_kernel(GfFwdArgs a) {
    constexpr uint32_t IDESC = umma_idesc_bf16(GF_BM, GF_BN);
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *sX = smem;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sX + GF_STAGES * GF_STAGE);
    uint64_t *a_full = bars;                       // [STAGES] loaders -> MMA
    uint64_t *a_empty = bars + GF_STAGES;          // [STAGES] MMA -> loaders
    uint64_t *t_full = bars + 2 * GF_STAGES;       // [2] MMA -> epilogue
    uint64_t *t_empty = t_full + 2;                // [2] epilogue -> MMA
    uint64_t *w_ready = t_empty + 2;               // [1] epilogue -> MMA: the weight block is in TMEM
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(w_ready + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x, cb = blockIdx.y, b = blockIdx.z;
    const int my_tiles = s < a.tiles ? (a.tiles - s + a.S - 1) / a.S : 0;

    if (threadIdx.x == 0) {
        for (int i = 0; i < GF_STAGES; ++i) { mbar_init(&a_full[i], GF_LOADERS); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4); }
        mbar_init(w_ready, 4);
        fence_barrier_init();
    }
    if (warp == GF_THREADS / 32 - 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // TMEM columns: [0, 128) weight heads, [128, 256) weight tails (two bf16 per column), [256, 384) / [384, 512) accumulators

    if (warp < GF_LOADERS / 32) {
        // ===================== loaders: fp32 rows of the cloud -> swizzled bf16 hi / lo tiles =====================
        const int tid = threadIdx.x;
        const int rsub = tid >> 4, c4 = tid & 15;
        const float *xb = a.x + (size_t)b * a.N * GF_K;
        int stage = 0;
        uint32_t phase = 0;
        const int items = my_tiles * GF_KCH;
        auto issue = [&](int item, float4 (&v)[8]) {
            const int n0 = (s + (item / GF_KCH) * a.S) * GF_BN, kc = item % GF_KCH;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int r = it * 16 + rsub;
                v[it] = n0 + r < a.N ? __ldg(reinterpret_cast<const float4 *>(xb + (size_t)(n0 + r) * GF_K + kc * GF_KB) + c4)
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        auto consume = [&](const float4 (&v)[8]) {
            mbar_wait(&a_empty[stage], phase ^ 1);
            uint8_t *hi_tile = sX + stage * GF_STAGE, *lo_tile = hi_tile + GF_TILE;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                uint2 hw, lw;
                gf_split4(v[it], hw, lw);
                const uint32_t off = gf_sw128(it * 16 + rsub, c4 * 4);
                *reinterpret_cast<uint2 *>(hi_tile + off) = hw;
                *reinterpret_cast<uint2 *>(lo_tile + off) = lw;
            }
            fence_proxy_async();
            mbar_arrive(&a_full[stage]);
            if (++stage == GF_STAGES) { stage = 0; phase ^= 1; }
        };
        float4 v0[8], v1[8], v2[8];
        if (items > 0) issue(0, v0);
        if (items > 1) issue(1, v1);
        for (int item = 0; item < items; item += 3) {
            if (item + 2 < items) issue(item + 2, v2);
            consume(v0);
            if (item + 1 >= items) break;
            if (item + 3 < items) issue(item + 3, v0);
            consume(v1);
            if (item + 2 >= items) break;
            if (item + 4 < items) issue(item + 4, v1);
            consume(v2);
        }
    } else if (warp == GF_THREADS / 32 - 1) {
        // ===================== MMA issuer (whole warp converged, one elected lane issues) =====================
        mbar_wait(w_ready, 0);
        tc_fence_after();
        const uint64_t b_desc0 = make_kmajor_sw128_desc(smem_u32(sX));
        int stage = 0, acc = 0;
        uint32_t phase = 0, accphase = 0;
        for (int t = 0; t < my_tiles; ++t) {
            mbar_wait(&t_empty[acc], accphase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + 256 + acc * GF_BN;
            for (int kc = 0; kc < GF_KCH; ++kc) {
                mbar_wait(&a_full[stage], phase);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint64_t b_desc = desc_advance(b_desc0, stage * GF_STAGE);
#pragma unroll
                    for (int ks = 0; ks < GF_KB / 16; ++ks) {
                        const uint32_t a_hi = tmem_base + kc * (GF_KB / 2) + ks * 8;
                        const uint32_t a_lo = a_hi + GF_K / 2;
                        const uint64_t b_hi = desc_advance(b_desc, ks * 32);
                        const uint64_t b_lo = desc_advance(b_desc, GF_TILE + ks * 32);
                        umma_bf16_ts(d_tmem, a_hi, b_hi, IDESC, (kc | ks) ? 1u : 0u);
                        umma_bf16_ts(d_tmem, a_hi, b_lo, IDESC, 1);
                        umma_bf16_ts(d_tmem, a_lo, b_hi, IDESC, 1);
                    }
                    umma_commit(&a_empty[stage]);
                    if (kc == GF_KCH - 1) umma_commit(&t_full[acc]);      // from the lane that issued the MMAs
                }
                __syncwarp();
                if (++stage == GF_STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; accphase ^= 1; }
        }
    } else {
        // ===================== epilogue: one output channel per thread =====================
        const int ew = warp & 3;                              // TMEM lane quarter (warps 8..11 -> 0..3)
        const int r = ew * 32 + lane;
        const int c = cb * GF_BM + r;
        // weight row -> TMEM (A operand): 64 channels at a time = 32 columns of heads + 32 columns of tails
        {
            const float4 *wr = reinterpret_cast<const float4 *>(a.w + (size_t)c * GF_K);
#pragma unroll 1
            for (int blk = 0; blk < GF_KCH; ++blk) {
                uint32_t hi[32], lo[32];
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    uint2 hw, lw;
                    gf_split4(__ldg(wr + blk * 16 + u), hw, lw);
                    hi[2 * u] = hw.x; hi[2 * u + 1] = hw.y;
                    lo[2 * u] = lw.x; lo[2 * u + 1] = lw.y;
                }
                tmem_st32(tmem_base + ((uint32_t)(ew * 32) << 16) + blk * 32, hi);
                tmem_st32(tmem_base + ((uint32_t)(ew * 32) << 16) + GF_K / 2 + blk * 32, lo);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(w_ready);
        }
        const float sg = __ldg(a.gamma + c) < 0.f ? -1.f : 1.f;
        float zmax = -CUDART_INF_F, s1 = 0.f, s2 = 0.f;
        int arg = 0;
        int acc = 0;
        uint32_t accphase = 0;
        for (int t = 0; t < my_tiles; ++t) {
            const int n0 = (s + t * a.S) * GF_BN;
            mbar_wait(&t_full[acc], accphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + 256 + acc * GF_BN;
            const int valid = min(GF_BN, a.N - n0);           // rows past N are zero-filled: they add 0 to the sums
#pragma unroll 1
            for (int ch = 0; ch < GF_BN / 32; ++ch) {
                uint32_t v[32];
                tmem_ld32(taddr + ch * 32, v);
                tmem_ld_wait();
                if (ch * 32 + 32 <= valid) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float y = __uint_as_float(v[j]);
                        const float z = sg * y;
                        if (z > zmax) { zmax = z; arg = n0 + ch * 32 + j; }
                        s1 += y;
                        s2 = fmaf(y, y, s2);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (ch * 32 + j < valid) {
                            const float y = __uint_as_float(v[j]);
                            const float z = sg * y;
                            if (z > zmax) { zmax = z; arg = n0 + ch * 32 + j; }
                            s1 += y;
                            s2 = fmaf(y, y, s2);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[acc]);
            if (++acc == 2) { acc = 0; accphase ^= 1; }
        }
        a.part[((size_t)b * a.S + s) * a.Cout + c] = make_float4(zmax, __int_as_float(arg), s1, s2);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == GF_THREADS / 32 - 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// per cloud: combine the slices in a fixed order, add the bias, GroupNorm statistics (fp64), x4 = relu(GN(extreme))
// saved: ysel [B][Cout], arg [B][Cout], sy [B][Cout] (sum over the points of y), stats [B][G][2] (mean, rstd)
__global__ void __launch_bounds__(256) gf_finalize_kernel(const float4 *__restrict__ part, const float *__restrict__ bias,
                                                          const float *__restrict__ gamma, const float *__restrict__ beta,
                                                          float *__restrict__ out, float *__restrict__ ysel, int *__restrict__ argn,
                                                          float *__restrict__ sy, float *__restrict__ stats, int N, int Cout, int S,
                                                          int G, float eps) {
    extern __shared__ double gsum[];                 // [2][G]
    const int b = blockIdx.x;
    const int cpg = Cout / G;
    for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) gsum[i] = 0.0;
    __syncthreads();
    for (int c = threadIdx.x; c < Cout; c += blockDim.x) {
        float zmax = -CUDART_INF_F;
        int arg = 0;
        double s1 = 0.0, s2 = 0.0;
        for (int s = 0; s < S; ++s) {
            const float4 p = part[((size_t)b * S + s) * Cout + c];
            if (p.x > zmax) { zmax = p.x; arg = __float_as_int(p.y); }
            s1 += (double)p.z;
            s2 += (double)p.w;
        }
        const float bs = bias ? bias[c] : 0.f;
        const float sg = gamma[c] < 0.f ? -1.f : 1.f;
        const size_t o = (size_t)b * Cout + c;
        ysel[o] = sg * zmax + bs;
        argn[o] = arg;
        const double t1 = s1 + (double)N * bs;
        const double t2 = s2 + 2.0 * bs * s1 + (double)N * bs * bs;
        sy[o] = (float)t1;
        atomicAdd(&gsum[c / cpg], t1);               // shared-memory fp64 atomics: order-dependent in the last bits only
        atomicAdd(&gsum[G + c / cpg], t2);
    }
    __syncthreads();
    const double cnt = (double)cpg * N;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        const double mean = gsum[g] / cnt;
        double var = gsum[G + g] / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        stats[((size_t)b * G + g) * 2] = (float)mean;
        stats[((size_t)b * G + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
    }
    __syncthreads();
    for (int c = threadIdx.x; c < Cout; c += blockDim.x) {
        const int g = c / cpg;
        const double mean = gsum[g] / cnt;
        double var = gsum[G + g] / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        const float rstd = (float)(1.0 / sqrt(var + (double)eps));
        const size_t o = (size_t)b * Cout + c;
        const float u = (ysel[o] - (float)mean) * rstd * gamma[c] + beta[c];
        out[o] = u > 0.f ? u : 0.f;
    }
}

// ------------------------------------------------------------------------------------------------ backward
// per cloud: du, s = rstd gamma du, and the GroupNorm coefficients (A_g, K_g)
__global__ void __launch_bounds__(256) gf_bwd_coef_kernel(const float *__restrict__ gout, const float *__restrict__ ysel,
                                                          const float *__restrict__ stats, const float *__restrict__ gamma,
                                                          const float *__restrict__ beta, float *__restrict__ sval,
                                                          float *__restrict__ du_out, float *__restrict__ duy_out,
                                                          float *__restrict__ coef, int N, int Cout, int G) {
    extern __shared__ double gsum[];                 // [2][G]
    const int b = blockIdx.x, cpg = Cout / G;
    for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) gsum[i] = 0.0;
    __syncthreads();
    for (int c = threadIdx.x; c < Cout; c += blockDim.x) {
        const int g = c / cpg;
        const float mean = stats[((size_t)b * G + g) * 2], rstd = stats[((size_t)b * G + g) * 2 + 1];
        const size_t o = (size_t)b * Cout + c;
        const float yh = (ysel[o] - mean) * rstd;
        const float u = yh * gamma[c] + beta[c];
        const float du = u > 0.f ? gout[o] : 0.f;
        sval[o] = rstd * gamma[c] * du;
        du_out[o] = du;
        duy_out[o] = du * yh;
        atomicAdd(&gsum[g], (double)gamma[c] * du);
        atomicAdd(&gsum[G + g], (double)gamma[c] * du * yh);
    }
    __syncthreads();
    const double cnt = (double)cpg * N;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        const double mean = stats[((size_t)b * G + g) * 2], rstd = stats[((size_t)b * G + g) * 2 + 1];
        const double m1 = gsum[g] / cnt, m2 = gsum[G + g] / cnt;
        coef[((size_t)b * G + g) * 2] = (float)(-rstd * m1 + rstd * rstd * m2 * mean);     // A_g
        coef[((size_t)b * G + g) * 2 + 1] = (float)(-rstd * rstd * m2);                    // K_g
    }
}

// dgamma, dbeta, dbias; cw[b][c] = A_bg + K_bg bias_c (weight of W_c in the constant row of dX and of Sx_b in dW)
__global__ void gf_bwd_affine_kernel(const float *__restrict__ sval, const float *__restrict__ du, const float *__restrict__ duy,
                                     const float *__restrict__ coef, const float *__restrict__ sy, const float *__restrict__ bias,
                                     float *__restrict__ dgamma, float *__restrict__ dbeta, float *__restrict__ dbias,
                                     float *__restrict__ cw, int B, int N, int Cout, int G) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= Cout) return;
    const int g = c / (Cout / G);
    const float bs = bias ? bias[c] : 0.f;
    double dg = 0.0, dbt = 0.0, dbs = 0.0;
    for (int b = 0; b < B; ++b) {
        const size_t o = (size_t)b * Cout + c;
        const float Ag = coef[((size_t)b * G + g) * 2], Kg = coef[((size_t)b * G + g) * 2 + 1];
        dg += duy[o];
        dbt += du[o];
        dbs += (double)sval[o] + (double)N * Ag + (double)Kg * sy[o];
        cw[o] = fmaf(Kg, bs, Ag);
    }
    dgamma[c] = (float)dg;
    dbeta[c] = (float)dbt;
    if (dbias) dbias[c] = (float)dbs;
}

// H[g][i][j] = sum_{c in g} W[c][i] W[c][j]     grid (K/16, K/16, G), block (16, 16)
__global__ void gf_bwd_h_kernel(const float *__restrict__ w, float *__restrict__ H, int Cout, int G) {
    const int g = blockIdx.z, cpg = Cout / G;
    const int j = blockIdx.x * 16 + threadIdx.x, i = blockIdx.y * 16 + threadIdx.y;
    const float *wg = w + (size_t)g * cpg * GF_K;
    float s = 0.f;
    for (int c = 0; c < cpg; ++c) s = fmaf(wg[(size_t)c * GF_K + i], wg[(size_t)c * GF_K + j], s);
    H[((size_t)g * GF_K + i) * GF_K + j] = s;
}

// M[b] = sum_g K_bg H[g]      grid (K*K/256, B)
__global__ void gf_bwd_m_kernel(const float *__restrict__ H, const float *__restrict__ coef, float *__restrict__ M, int G) {
    const int b = blockIdx.y, e = blockIdx.x * 256 + threadIdx.x;
    float s = 0.f;
    for (int g = 0; g < G; ++g) s = fmaf(coef[((size_t)b * G + g) * 2 + 1], H[(size_t)g * GF_K * GF_K + e], s);
    M[(size_t)b * GF_K * GF_K + e] = s;
}

// cvec[b][k] = sum_c cw[b][c] W[c][k]      grid B, block K
__global__ void gf_bwd_cvec_kernel(const float *__restrict__ cw, const float *__restrict__ w, float *__restrict__ cvec, int Cout) {
    const int b = blockIdx.x, k = threadIdx.x;
    float s = 0.f;
    for (int c = 0; c < Cout; ++c) s = fmaf(cw[(size_t)b * Cout + c], w[(size_t)c * GF_K + k], s);
    cvec[(size_t)b * GF_K + k] = s;
}

// dX[b][n*_c][:] += s_bc W_c     one warp per (b, c): 256 floats = two float4 vector reductions per lane
__global__ void gf_bwd_sparse_kernel(const float *__restrict__ sval, const int *__restrict__ argn, const float *__restrict__ w,
                                     float *__restrict__ dx, int N, int Cout, int total) {
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (gw >= total) return;
    const int b = gw / Cout, c = gw % Cout;
    const float s = sval[gw];
    if (s == 0.f) return;
    float *row = dx + ((size_t)b * N + argn[gw]) * GF_K;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float4 wv = __ldg(reinterpret_cast<const float4 *>(w + (size_t)c * GF_K) + h * 32 + lane);
        atomicAdd(reinterpret_cast<float4 *>(row) + h * 32 + lane, make_float4(s * wv.x, s * wv.y, s * wv.z, s * wv.w));
    }
}

// sxp[b][p][k] = sum of x over the p-th slice of the cloud's points     grid (P, B), block K
__global__ void gf_bwd_colsum_kernel(const float *__restrict__ x, float *__restrict__ sxp, int N, int P) {
    const int b = blockIdx.y, p = blockIdx.x, k = threadIdx.x;
    const int per = (N + P - 1) / P, lo = p * per, hi = min(N, lo + per);
    float s = 0.f;
    for (int n = lo; n < hi; ++n) s += x[((size_t)b * N + n) * GF_K + k];
    sxp[((size_t)b * P + p) * GF_K + k] = s;
}

// gs[g][i][j] = sum_b K_bg G_b[i][j], G_b = sum over the splits of the four 64-row partial blocks   grid (K*K/256, G)
__global__ void gf_bwd_gsum_kernel(const float *__restrict__ gpart, const float *__restrict__ coef, float *__restrict__ gs, int B,
                                   int G, int splits) {
    const int g = blockIdx.y, e = blockIdx.x * 256 + threadIdx.x;
    const int i = e / GF_K, j = e % GF_K;
    const int blk = i / 64, m = i % 64;
    double s = 0.0;
    for (int b = 0; b < B; ++b) {
        const float kg = coef[((size_t)b * G + g) * 2 + 1];
        float gb = 0.f;
        for (int sp = 0; sp < splits; ++sp)
            gb += gpart[((((size_t)blk * B + b) * splits + sp) * 64 + m) * GF_K + j];
        s += (double)kg * gb;
    }
    gs[(size_t)g * GF_K * GF_K + e] = (float)s;
}

// dW[c][k] = sum_b s_bc x_b[n*_bc][k]  +  sum_b cw_bc Sx_b[k]  +  sum_i W[c][i] gs[g(c)][i][k]      grid Cout, block K
__global__ void gf_bwd_dw_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ sval,
                                 const int *__restrict__ argn, const float *__restrict__ cw, const float *__restrict__ sxp,
                                 const float *__restrict__ gs, float *__restrict__ dw, int B, int N, int Cout, int G, int P) {
    __shared__ float wrow[GF_K];
    const int c = blockIdx.x, k = threadIdx.x;
    const int g = c / (Cout / G);
    wrow[k] = w[(size_t)c * GF_K + k];
    __syncthreads();
    float t3 = 0.f;
    const float *gg = gs + (size_t)g * GF_K * GF_K;
#pragma unroll 8
    for (int i = 0; i < GF_K; ++i) t3 = fmaf(wrow[i], gg[(size_t)i * GF_K + k], t3);
    double acc = (double)t3;
    for (int b = 0; b < B; ++b) {
        const size_t o = (size_t)b * Cout + c;
        float sx = 0.f;
        for (int p = 0; p < P; ++p) sx += sxp[((size_t)b * P + p) * GF_K + k];
        acc += (double)sval[o] * x[((size_t)b * N + argn[o]) * GF_K + k] + (double)cw[o] * sx;
    }
    dw[(size_t)c * GF_K + k] = (float)acc;
}

// ------------------------------------------------------------------------------------------------ host side
constexpr int GF_COLSUM_P = 32;

struct GfSaved { float *ysel, *sy, *stats; int *argn; };
struct GfWs {
    float4 *part;                                   // forward
    float *sval, *du, *duy, *coef, *cw, *H, *M, *cvec, *sxp, *gpart, *gs;   // backward
};

static int gf_slices(int B, int Cout, int tiles) {
    // enough CTAs for ~3 waves of 148 SMs, at least ~8 tiles each
    int S = ceil_div(3 * kNumSMs, B * (Cout / GF_BM));
    if (S > ceil_div(tiles, 8)) S = ceil_div(tiles, 8);
    return S < 1 ? 1 : S;
}

static int gf_gram_splits(int N) {
    const int rows = ceil_div(ceil_div(N, GF_GRAM_SPLITS), 64) * 64;
    return ceil_div(N, rows);
}

static int gf_check(const gcanet_global_feature_desc *d) {
    GCANET_REQUIRE(d != nullptr, "global_feature: null descriptor");
    GCANET_REQUIRE(d->B >= 1 && d->B <= 65535 && d->N >= 1, "global_feature: bad shape B=%d N=%d", d->B, d->N);
    GCANET_REQUIRE(d->K == GF_K, "global_feature: K=%d (this build covers K = 256 = 64 + 64 + 128, M4:507)", d->K);
    GCANET_REQUIRE(d->Cout >= GF_BM && d->Cout % GF_BM == 0 && d->Cout <= 4096, "global_feature: Cout=%d must be a multiple of 128", d->Cout);
    GCANET_REQUIRE(d->groups >= 1 && d->groups <= 64 && d->Cout % d->groups == 0, "global_feature: groups=%d must divide Cout", d->groups);
    GCANET_REQUIRE(d->eps > 0.f, "global_feature: eps must be positive");
    GCANET_REQUIRE((long long)d->B * d->N < 2147483647ll, "global_feature: B * N does not fit 32 bits");
    return GCANET_OK;
}

static size_t gf_plan_saved(const gcanet_global_feature_desc *d, void *base, GfSaved *s) {
    Carver cv(base);
    const size_t bc = (size_t)d->B * d->Cout;
    float *ysel = cv.take<float>(bc);
    float *sy = cv.take<float>(bc);
    float *stats = cv.take<float>((size_t)d->B * d->groups * 2);
    int *argn = cv.take<int>(bc);
    if (s) { s->ysel = ysel; s->sy = sy; s->stats = stats; s->argn = argn; }
    return cv.off;
}

static size_t gf_plan_ws(const gcanet_global_feature_desc *d, void *base, GfWs *w) {
    Carver cv(base);
    const size_t bc = (size_t)d->B * d->Cout;
    const int tiles = ceil_div(d->N, GF_BN);
    float4 *part = cv.take<float4>((size_t)d->B * gf_slices(d->B, d->Cout, tiles) * d->Cout);
    float *sval = cv.take<float>(bc), *du = cv.take<float>(bc), *duy = cv.take<float>(bc), *cw = cv.take<float>(bc);
    float *coef = cv.take<float>((size_t)d->B * d->groups * 2);
    float *H = cv.take<float>((size_t)d->groups * GF_K * GF_K);
    float *M = cv.take<float>((size_t)d->B * GF_K * GF_K);
    float *cvec = cv.take<float>((size_t)d->B * GF_K);
    float *sxp = cv.take<float>((size_t)d->B * GF_COLSUM_P * GF_K);
    float *gpart = cv.take<float>((size_t)4 * d->B * gf_gram_splits(d->N) * 64 * GF_K);
    float *gs = cv.take<float>((size_t)d->groups * GF_K * GF_K);
    if (w) { w->part = part; w->sval = sval; w->du = du; w->duy = duy; w->coef = coef; w->cw = cw; w->H = H; w->M = M; w->cvec = cvec;
             w->sxp = sxp; w->gpart = gpart; w->gs = gs; }
    return cv.off;
}

}  // namespace gcanet

using namespace gcanet;

extern "C" size_t gcanet_global_feature_saved_bytes(const gcanet_global_feature_desc *d) {
    if (gf_check(d) != GCANET_OK) return 0;
    return gf_plan_saved(d, nullptr, nullptr);
}

extern "C" size_t gcanet_global_feature_workspace_bytes(const gcanet_global_feature_desc *d) {
    if (gf_check(d) != GCANET_OK) return 0;
    return gf_plan_ws(d, nullptr, nullptr);
}

extern "C" int gcanet_global_feature_forward(const gcanet_global_feature_desc *d, const float *x_nc, const float *weight,
                                             const float *bias, const float *gamma, const float *beta, float *out, void *saved,
                                             void *ws, size_t ws_bytes, gcanet_stream_t stream) {
    int rc = gf_check(d);
    if (rc) return rc;
    GCANET_REQUIRE(x_nc && weight && gamma && beta && out && saved, "global_feature_forward: null pointer");
    GCANET_REQUIRE(((reinterpret_cast<uintptr_t>(x_nc) | reinterpret_cast<uintptr_t>(weight)) & 15) == 0,
                   "global_feature_forward: x and weight must be 16-byte aligned");
    const size_t need = gf_plan_ws(d, nullptr, nullptr);
    if (ws == nullptr || ws_bytes < need || reinterpret_cast<uintptr_t>(ws) % kAlign || reinterpret_cast<uintptr_t>(saved) % kAlign) {
        set_error("global_feature_forward: workspace too small or misaligned (%zu given, %zu needed)", ws_bytes, need);
        return GCANET_ERR_WORKSPACE;
    }
    GfSaved sv; GfWs w;
    gf_plan_saved(d, saved, &sv);
    gf_plan_ws(d, ws, &w);
    cudaStream_t st = as_stream(stream);
    const int tiles = ceil_div(d->N, GF_BN);
    const int S = gf_slices(d->B, d->Cout, tiles);
    const size_t smem = 1024 + (size_t)GF_STAGES * GF_STAGE + 256;
    GCANET_CUDA_OK(cudaFuncSetAttribute(gf_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GfFwdArgs fa{x_nc, weight, gamma, w.part, d->N, d->Cout, S, tiles};
    gf_forward_kernel<<<dim3(S, d->Cout / GF_BM, d->B), GF_THREADS, smem, st>>>(fa);
    GCANET_LAUNCH_OK("gf_forward_kernel");
    gf_finalize_kernel<<<d->B, 256, 2 * d->groups * sizeof(double), st>>>(w.part, bias, gamma, beta, out, sv.ysel, sv.argn, sv.sy, sv.stats,
                                                                         d->N, d->Cout, S, d->groups, d->eps);
    GCANET_LAUNCH_OK("gf_finalize_kernel");
    return GCANET_OK;
}

extern "C" int gcanet_global_feature_backward(const gcanet_global_feature_desc *d, const float *x_nc, const float *weight,
                                              const float *bias, const float *gamma, const float *beta, const float *grad_out,
                                              const void *saved, float *grad_x_nc, float *grad_weight, float *grad_bias,
                                              float *grad_gamma, float *grad_beta, void *ws, size_t ws_bytes, gcanet_stream_t stream) {
    int rc = gf_check(d);
    if (rc) return rc;
    GCANET_REQUIRE(x_nc && weight && gamma && beta && grad_out && saved && grad_weight && grad_gamma && grad_beta,
                   "global_feature_backward: null pointer");
    const size_t need = gf_plan_ws(d, nullptr, nullptr);
    if (ws == nullptr || ws_bytes < need || reinterpret_cast<uintptr_t>(ws) % kAlign) {
        set_error("global_feature_backward: workspace too small or misaligned (%zu given, %zu needed)", ws_bytes, need);
        return GCANET_ERR_WORKSPACE;
    }
    GfSaved sv; GfWs w;
    gf_plan_saved(d, const_cast<void *>(saved), &sv);
    gf_plan_ws(d, ws, &w);
    cudaStream_t st = as_stream(stream);
    const int B = d->B, N = d->N, Cout = d->Cout, G = d->groups;
    gf_bwd_coef_kernel<<<B, 256, 2 * G * sizeof(double), st>>>(grad_out, sv.ysel, sv.stats, gamma, beta, w.sval, w.du, w.duy, w.coef, N, Cout, G);
    GCANET_LAUNCH_OK("gf_bwd_coef_kernel");
    gf_bwd_affine_kernel<<<ceil_div(Cout, 128), 128, 0, st>>>(w.sval, w.du, w.duy, w.coef, sv.sy, bias, grad_gamma, grad_beta, grad_bias,
                                                              w.cw, B, N, Cout, G);
    GCANET_LAUNCH_OK("gf_bwd_affine_kernel");
    // weight gradient: Gram matrices of the clouds (tensor cores), column sums, then one pass per output channel
    const int splits = gf_gram_splits(N);
    for (int blk = 0; blk < 4; ++blk) {
        rc = gemm_tn_tc_batched(x_nc + blk * 64, GF_K, (long long)N * GF_K, x_nc, GF_K, (long long)N * GF_K,
                                w.gpart + (size_t)blk * B * splits * 64 * GF_K, N, splits, B, st);
        if (rc > 0) { set_error("global_feature_backward: Gram product not covered (N=%d)", N); return GCANET_ERR_INVALID_ARGUMENT; }
        if (rc) return rc;
    }
    gf_bwd_gsum_kernel<<<dim3(GF_K * GF_K / 256, G), 256, 0, st>>>(w.gpart, w.coef, w.gs, B, G, splits);
    GCANET_LAUNCH_OK("gf_bwd_gsum_kernel");
    gf_bwd_colsum_kernel<<<dim3(GF_COLSUM_P, B), GF_K, 0, st>>>(x_nc, w.sxp, N, GF_COLSUM_P);
    GCANET_LAUNCH_OK("gf_bwd_colsum_kernel");
    gf_bwd_dw_kernel<<<Cout, GF_K, 0, st>>>(x_nc, weight, w.sval, sv.argn, w.cw, w.sxp, w.gs, grad_weight, B, N, Cout, G, GF_COLSUM_P);
    GCANET_LAUNCH_OK("gf_bwd_dw_kernel");
    if (grad_x_nc) {
        gf_bwd_h_kernel<<<dim3(GF_K / 16, GF_K / 16, G), dim3(16, 16), 0, st>>>(weight, w.H, Cout, G);
        GCANET_LAUNCH_OK("gf_bwd_h_kernel");
        gf_bwd_m_kernel<<<dim3(GF_K * GF_K / 256, B), 256, 0, st>>>(w.H, w.coef, w.M, G);
        GCANET_LAUNCH_OK("gf_bwd_m_kernel");
        gf_bwd_cvec_kernel<<<B, GF_K, 0, st>>>(w.cw, weight, w.cvec, Cout);
        GCANET_LAUNCH_OK("gf_bwd_cvec_kernel");
        // dX[b] = X[b] M_b + c_b  (M_b symmetric: the "transposed weight" the GEMM wants is M_b itself), 128 columns per launch
        for (int half = 0; half < 2; ++half) {
            rc = gemm_tc_batched_try(x_nc, GF_K, (long long)N * GF_K, w.M + (size_t)half * 128 * GF_K, GF_K, (long long)GF_K * GF_K,
                                     grad_x_nc + half * 128, GF_K, (long long)N * GF_K, w.cvec + half * 128, GF_K, N, 128, GF_K, B, st);
            if (rc > 0) { set_error("global_feature_backward: dX product not covered"); return GCANET_ERR_INVALID_ARGUMENT; }
            if (rc) return rc;
        }
        const int total = B * Cout;
        gf_bwd_sparse_kernel<<<ceil_div(total, 8), 256, 0, st>>>(w.sval, sv.argn, weight, grad_x_nc, N, Cout, total);
        GCANET_LAUNCH_OK("gf_bwd_sparse_kernel");
    }
    return GCANET_OK;
}
