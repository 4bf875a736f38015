// Feature-space kNN (C = 64 / 128, L2 metric) on the 5th-generation tensor cores.
//
//   prep      x[B][C][N] fp32 -> xs[B][N][2C] bf16 = (hi | lo) with hi = bf16(x), lo = bf16(x - hi),
//             x_nc[B][N][C] fp32 (point-major, for the exact re-rank), |x|^2 (reference order)
//   scan      approximate distances d~ = |k|^2 - 2 q.k for 128-query x 64-key tiles:
//               TMA producer warp: query tile once, then key tiles through an mbarrier ring
//               tcgen05.mma issuer: D (TMEM, fp32) = Qhi Khi^T + Qhi Klo^T + Qlo Khi^T
//                         (bf16 x 3 split, |error| <= ~2^-14 |q||k|), accumulators ring-buffered in TMEM
//               epilogue warps: tcgen05.ld accumulator rows, threshold filter, survivors appended to the row's
//                         candidate list.  Every key whose approximate distance is within `margin` (>= 2 x the
//                         error bound) of the approximate k-th survives, so the exact k nearest are always
//                         among the candidates.
//             Two kernels:
//               knn_tcp_scan_kernel (N >= 1024, k <= 128; the default): the cloud is sorted along a Morton curve in
//                         the space of its three leading principal directions, every 64-key tile has a bounding
//                         box there, and a query tile walks the key tiles nearest-box-first and stops when no
//                         row's bound can be beaten.  Pass A finds each row's threshold from per-column-slot
//                         minima held in registers, pass B collects the candidates below it.
//               knn_tc_scan_kernel (clouds under 1024 points, or GCANET_KNN_FLAG_NO_PRUNE): every tile once, in a
//                         strided order, streaming threshold with warp-cooperative list compaction.
//   rerank    one warp per query: exact fp32 distances of the candidates whose membership is in doubt, in the
//             reference's expansion arithmetic, ranked by (distance, index).
//   fallback  rows whose candidate list overflowed (massive ties) are listed and redone by the CUDA-core scan.
//
// No N x N matrix and no approximate distance ever decides the result: tensor cores only prune.
#include "common.cuh"
#include "tc_ptx.cuh"

#include <cuda.h>
#include <cub/block/block_radix_sort.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <cuda_bf16.h>
#include <math_constants.h>
#include <stdlib.h>
#include <string.h>

namespace gcanet {

constexpr unsigned FULLW = 0xffffffffu;
constexpr int TC_BM = 128;            // queries per CTA  (UMMA M)
constexpr int TC_BN = 64;             // keys per tile    (UMMA N); 64 keeps a CTA at ~66 KB so three share an SM
constexpr int TC_KB = 64;             // bf16 elements per 128-byte swizzle row
constexpr int TC_CAP = 256;           // candidate list capacity per query (full scan)
constexpr int TCP_CAP = 512;          // ... on the box-pruned path: dense feature regions hold hundreds of keys within the
                                      // margin of the k-th distance, and they must not overflow into the slow fallback
constexpr int TC_SLACK = 8;           // bisection stops once the bound keeps <= k + slack entries
constexpr int TC_THREADS = 192;       // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue
constexpr int TC_NRING = 8;           // key-norm ring slots (producer is never more than 4 tiles ahead of the epilogue)
constexpr float TC_MARGIN = 7.0e-4f;  // ~2^-10.5 : margin = TC_MARGIN * |q| * max|k|  

// kind::f16, A = B = bf16, D = fp32, both K-major, M = 128, N = 128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

// ---------------------------------------------------------------------------------
// prep
// ---------------------------------------------------------------------------------
// xs[b][n][0:C] = hi, xs[b][n][C:2C] = lo ; x_nc[b][n][c] = x ; transposes through smem
// `inv` (may be null): xs row of point n is written at sorted position inv[b][n] (pruned path); x_nc keeps
// the original order either way
__global__ void tc_prep_kernel(const float *__restrict__ x, __nv_bfloat16 *__restrict__ xs, float *__restrict__ x_nc,
                               const int *__restrict__ inv, int C, int N) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const float *s = x + (size_t)b * C * N;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int c = c0 + r, n = n0 + threadIdx.x;
        tile[r][threadIdx.x] = (c < C && n < N) ? s[(size_t)c * N + n] : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int n = n0 + r, c = c0 + threadIdx.x;
        if (n < N && c < C) {
            float v = tile[threadIdx.x][r];
            __nv_bfloat16 hi = __float2bfloat16_rn(v);
            __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
            const size_t row = (size_t)b * N + n;
            const size_t srow = inv ? (size_t)b * N + inv[row] : row;
            xs[srow * 2 * C + c] = hi;
            xs[srow * 2 * C + C + c] = lo;
            x_nc[row * C + c] = v;
        }
    }
}

// Same, 64 channels x 64 points per CTA with 16-byte loads and 16 / 8-byte stores (N % 4 == 0, C % 64 == 0, 16-byte
// aligned bases); identical values.  The 2-byte stores of the kernel above left it at a third of the HBM rate.
__global__ void __launch_bounds__(256) tc_prep_wide_kernel(const float *__restrict__ x, __nv_bfloat16 *__restrict__ xs,
                                                           float *__restrict__ x_nc, const int *__restrict__ inv, int C, int N) {
    __shared__ float tile[64][65];
    const int b = blockIdx.z, n0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
    const float *s = x + (size_t)b * C * N;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int e = threadIdx.x + 256 * i, c = e >> 4, n = (e & 15) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n0 + n < N) v = __ldg(reinterpret_cast<const float4 *>(s + (size_t)(c0 + c) * N + n0 + n));
        tile[c][n] = v.x; tile[c][n + 1] = v.y; tile[c][n + 2] = v.z; tile[c][n + 3] = v.w;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int e = threadIdx.x + 256 * i, n = e >> 4, c = (e & 15) * 4;
        if (n0 + n >= N) continue;
        const size_t row = (size_t)b * N + n0 + n;
        const size_t srow = inv ? (size_t)b * N + inv[row] : row;
        float v[4];
        __nv_bfloat16 hi[4], lo[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            v[u] = tile[c + u][n];
            hi[u] = __float2bfloat16_rn(v[u]);
            lo[u] = __float2bfloat16_rn(v[u] - __bfloat162float(hi[u]));
        }
        *reinterpret_cast<float4 *>(x_nc + row * C + c0 + c) = make_float4(v[0], v[1], v[2], v[3]);
        uint2 ph, pl;
        ph.x = (unsigned)__bfloat16_as_ushort(hi[0]) | ((unsigned)__bfloat16_as_ushort(hi[1]) << 16);
        ph.y = (unsigned)__bfloat16_as_ushort(hi[2]) | ((unsigned)__bfloat16_as_ushort(hi[3]) << 16);
        pl.x = (unsigned)__bfloat16_as_ushort(lo[0]) | ((unsigned)__bfloat16_as_ushort(lo[1]) << 16);
        pl.y = (unsigned)__bfloat16_as_ushort(lo[2]) | ((unsigned)__bfloat16_as_ushort(lo[3]) << 16);
        *reinterpret_cast<uint2 *>(xs + srow * 2 * C + c0 + c) = ph;
        *reinterpret_cast<uint2 *>(xs + srow * 2 * C + C + c0 + c) = pl;
    }
}

// nmax[b] = max_n norm[b][n]; norm_pad[b][0:Npad] = norm[b] followed by +inf   (one CTA per cloud)
__global__ void tc_normmax_kernel(const float *__restrict__ norm, float *__restrict__ nmax,
                                  float *__restrict__ norm_pad, int N, int Npad) {
    __shared__ float red[32];
    const int b = blockIdx.x;
    float m = 0.f;
    for (int n = threadIdx.x; n < Npad; n += blockDim.x) {
        float v = n < N ? norm[(size_t)b * N + n] : CUDART_INF_F;
        if (norm_pad) norm_pad[(size_t)b * Npad + n] = v;
        if (n < N) m = fmaxf(m, v);
    }
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULLW, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULLW, m, o));
        if (threadIdx.x == 0) nmax[b] = m;
    }
}

// ---------------------------------------------------------------------------------
// scan
// ---------------------------------------------------------------------------------
struct TcScanArgs {
    const float *norm_pad;  // [B][Npad] key norms, Npad = tiles * TC_BN, +inf past N (masks the zero-filled keys)
    int Npad;
    const float *norm;   // [B][N]
    const float *nmax;   // [B]
    uint2 *cand;         // [B][N][TC_CAP]  (approx distance bits, key index)
    int *cand_cnt;       // [B][N]
    int *overflow;       // [B][N]   1 = list overflowed, row must be redone by the fallback
    int N, k, tiles;     // tiles = ceil(N / TC_BN)
    int debug_no_append; // measurement aid (GCANET_TC_DEBUG=1): thresholds start at -inf, nothing is ever appended
    int tile_stride;     // key tiles are visited as (t * tile_stride) % tiles, stride coprime to tiles:
                         // a spatially sorted cloud then looks like a random stream to the thresholds
    // hybrid launch next to the box-pruned kernel (which takes the clouds with structured[b] != 0):
    const int *structured;  // null, or [B]: this kernel only takes the clouds with structured[b] == 0
    int cap;                // row stride of `cand` in entries (TC_CAP, or TCP_CAP in the hybrid launch)
    int split_counts;       // 1: counts are written as (cnt, 0) pairs at cand_cnt[2 row], the pruned path's layout
};

// Bisection over a warp-distributed list (entries e = s*32 + lane, +inf padding): returns hi with
// count(d <= hi) >= k, stopping once <= k + TC_SLACK entries qualify or the interval is exhausted.
template <int SL>
__device__ __forceinline__ float select_bound(const float (&dv)[SL], int n, int k, float *lo_strict = nullptr) {
    float mn = CUDART_INF_F, mx = -CUDART_INF_F;
#pragma unroll
    for (int s = 0; s < SL; ++s) {
        mn = fminf(mn, dv[s]);
        if (dv[s] < CUDART_INF_F) mx = fmaxf(mx, dv[s]);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(FULLW, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(FULLW, mx, o));
    }
    float lo = mn, hi = mx, lo_s = -CUDART_INF_F;     // lo_s: count(d <= lo_s) < k is PROVEN
    int c_lo = 0;
    if (n > k) {
        int c_hi = n;
        for (int it = 0; it < 24 && c_hi > k + TC_SLACK; ++it) {
            float mid = 0.5f * lo + 0.5f * hi;
            if (!(mid > lo && mid < hi)) break;       // interval exhausted (ties)
            int c = 0;
#pragma unroll
            for (int s = 0; s < SL; ++s) c += (dv[s] <= mid) ? 1 : 0;
            c = __reduce_add_sync(FULLW, c);
            if (c >= k) { hi = mid; c_hi = c; } else { lo = mid; lo_s = mid; c_lo = c; }
        }
        if (lo_strict) {
            // tighten the proven lower side too, so the ambiguous band (lo_s, hi] holds ~2*slack entries
            float h2 = hi;
            for (int it = 0; it < 12 && c_lo < k - TC_SLACK; ++it) {
                float mid = 0.5f * lo + 0.5f * h2;
                if (!(mid > lo && mid < h2)) break;
                int c = 0;
#pragma unroll
                for (int s = 0; s < SL; ++s) c += (dv[s] <= mid) ? 1 : 0;
                c = __reduce_add_sync(FULLW, c);
                if (c < k) { lo = mid; lo_s = mid; c_lo = c; } else { h2 = mid; }
            }
        }
    }
    if (lo_strict) *lo_strict = lo_s;
    return hi;
}

// Warp-cooperative compaction of one row's candidate list (entries [0, n) at `ptr`): keeps
// d <= bound + margin, returns the new count and threshold.  `ptr`, `n`, `margin` are warp-uniform.
template <int CAPACITY = TC_CAP>
__device__ __forceinline__ void compact_row(uint2 *ptr, int n, int k, float margin, int lane, int &new_cnt, float &new_thr) {
    constexpr int SL = CAPACITY / 32;
    float dv[SL];
    uint32_t di[SL];
#pragma unroll
    for (int s = 0; s < SL; ++s) {
        int e = s * 32 + lane;
        dv[s] = CUDART_INF_F;
        di[s] = 0;
        if (e < n) {
            uint2 t = ptr[e];
            dv[s] = __uint_as_float(t.x);
            di[s] = t.y;
        }
    }
    const float keep_below = select_bound<SL>(dv, n, k) + margin;
    __syncwarp();
    int base = 0;
#pragma unroll
    for (int s = 0; s < SL; ++s) {
        bool keep = dv[s] <= keep_below;          // padding (+inf) never kept
        unsigned m = __ballot_sync(FULLW, keep);
        if (keep) ptr[base + __popc(m & ((1u << lane) - 1))] = make_uint2(__float_as_uint(dv[s]), di[s]);
        base += __popc(m);
    }
    __syncwarp();
    new_cnt = base;
    new_thr = keep_below;
}

// MODE 0 = product.  Measurement aids (GCANET_TC_DEBUG=1/2/3, tools/time_knn.py), separate instantiations so the
// product kernel's code is untouched: 1 = appends off (thresholds at -inf), 2 = TMEM loads + a trivial
// consumer (no distance, no compare), 3 = no TMEM loads at all (TMA + MMA + barrier hand-offs only).
template <int C, int MODE>
__global__ void __launch_bounds__(TC_THREADS, C == 64 ? 3 : 1)
knn_tc_scan_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k, TcScanArgs a) {
    constexpr int NBLK = 2 * C / TC_KB;                 // 128-byte K blocks per row: hi blocks then lo blocks
    constexpr int NH = C / TC_KB;                       // hi (= lo) blocks
    constexpr int ABLK_BYTES = TC_BM * 128;             // one K block of the 128-query tile: 16 KB
    constexpr int A_BYTES = NBLK * ABLK_BYTES;          // 32 KB (C=64) / 64 KB (C=128)
    constexpr int BLK_BYTES = TC_BN * 128;              // one K block of a 64-key tile: 8 KB
    constexpr int TILE_BYTES = NBLK * BLK_BYTES;        // 16 KB (C=64) / 32 KB (C=128)
    constexpr int STAGES = 2;                           // C = 64: 32 + 2*16 KB per CTA -> three CTAs share an SM
    constexpr int ACC = 2;                              // TMEM accumulator stages

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // align inside the shared window with plain pointer arithmetic: casting through an integer would
    // make every later access a generic load/store instead of LDS/STS
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *sA = smem;                                  // query tile
    uint8_t *sB = smem + A_BYTES;                        // STAGES key tiles
    float *s_rn = reinterpret_cast<float *>(sB + STAGES * TILE_BYTES);      // [TC_NRING][TC_BN] key norms
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_rn + TC_NRING * TC_BN);
    uint64_t *full = bars;                 // [STAGES]  TMA -> MMA
    uint64_t *empty = bars + STAGES;       // [STAGES]  MMA -> TMA
    uint64_t *a_full = bars + 2 * STAGES;  // [1]
    uint64_t *t_full = a_full + 1;         // [ACC]     MMA -> epilogue
    uint64_t *t_empty = t_full + ACC;      // [ACC]     epilogue -> MMA
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(t_empty + ACC);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int q0 = blockIdx.x * TC_BM;
    const int tiles = a.tiles;
    if (a.structured != nullptr && a.structured[b] != 0) return;     // the box-pruned kernel has this cloud

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_k);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(a_full, 1);
        for (int s = 0; s < ACC; ++s) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, ACC * TC_BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_expect_tx(a_full, A_BYTES);
#pragma unroll
            for (int kb = 0; kb < NBLK; ++kb) tma_load_3d(sA + kb * ABLK_BYTES, &tmap_q, a_full, kb * TC_KB, q0, b);
            const float *rn_g = a.norm_pad + (size_t)b * a.Npad;
            int stage = 0;
            uint32_t phase = 0;
            for (int t = 0; t < tiles; ++t) {
                mbar_wait_backoff(&empty[stage], phase ^ 1);
                mbar_expect_tx(&full[stage], TILE_BYTES + TC_BN * sizeof(float));
                uint8_t *dst = sB + stage * TILE_BYTES;
                const int kt = (int)(((long long)t * a.tile_stride) % tiles);
#pragma unroll
                for (int kb = 0; kb < NBLK; ++kb) tma_load_3d(dst + kb * BLK_BYTES, &tmap_k, &full[stage], kb * TC_KB, kt * TC_BN, b);
                // this tile's key norms ride on the same barrier; by the time the MMA that consumed the
                // stage has committed to t_full, the epilogue may read them
                bulk_load_1d(s_rn + (t % TC_NRING) * TC_BN, rn_g + (size_t)kt * TC_BN, TC_BN * sizeof(float), &full[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // whole warp converged, one elected lane issues (tc_ptx.cuh: elect_one_sync), descriptors advanced from a base
        {
            mbar_wait(a_full, 0);
            tc_fence_after();
            const uint64_t a_desc0 = make_kmajor_sw128_desc(smem_u32(sA));
            const uint64_t b_desc0 = make_kmajor_sw128_desc(smem_u32(sB));
            int stage = 0, acc = 0;
            uint32_t phase = 0, accphase = 0;
            for (int t = 0; t < tiles; ++t) {
                mbar_wait_backoff(&t_empty[acc], accphase ^ 1);
                mbar_wait_backoff(&full[stage], phase);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint64_t b_desc = desc_advance(b_desc0, stage * TILE_BYTES);
                    const uint32_t d_tmem = tmem_base + acc * TC_BN;
#pragma unroll
                    for (int hb = 0; hb < NH; ++hb) {
#pragma unroll
                        for (int ks = 0; ks < TC_KB / 16; ++ks) {
                            const uint32_t koff = ks * 32;        // 16 bf16 = 32 bytes inside the swizzled row
                            const uint64_t a_hi = desc_advance(a_desc0, hb * ABLK_BYTES + koff);
                            const uint64_t a_lo = desc_advance(a_desc0, (NH + hb) * ABLK_BYTES + koff);
                            const uint64_t b_hi = desc_advance(b_desc, hb * BLK_BYTES + koff);
                            const uint64_t b_lo = desc_advance(b_desc, (NH + hb) * BLK_BYTES + koff);
                            umma_bf16(d_tmem, a_hi, b_hi, kIdesc, (hb | ks) ? 1u : 0u);
                            umma_bf16(d_tmem, a_hi, b_lo, kIdesc, 1);
                            umma_bf16(d_tmem, a_lo, b_hi, kIdesc, 1);
                        }
                    }
                    umma_commit(&empty[stage]);      // smem slot reusable once these MMAs have read it
                    umma_commit(&t_full[acc]);       // accumulator ready for the epilogue
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
                if (++acc == ACC) { acc = 0; accphase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue: one query row per thread =====================
        const int ew = warp & 3;                          // TMEM lane group this warp may access
        const int row = ew * 32 + lane;                   // accumulator row = TMEM lane
        const int q = q0 + row;
        const bool active = q < a.N;
        const size_t grow = (size_t)b * a.N + (active ? q : 0);
        uint2 *buf = a.cand + grow * a.cap;
        const float qn = active ? a.norm[grow] : 0.f;
        const float margin = TC_MARGIN * sqrtf(qn * a.nmax[b]);
        float thr = (active && !a.debug_no_append) ? CUDART_INF_F : -CUDART_INF_F;   // inactive rows never append
        int cnt = 0;
        bool ovf = false;

        int acc = 0;
        uint32_t accphase = 0;
        for (int t = 0; t < tiles; ++t) {
            const int kt = (int)(((long long)t * a.tile_stride) % tiles);
            mbar_wait(&t_full[acc], accphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * TC_BN;
            if constexpr (MODE >= 2) {
                if constexpr (MODE == 2) {
                    uint32_t w[2][32];
                    tmem_ld32(taddr, w[0]);
#pragma unroll
                    for (int ch = 0; ch < TC_BN / 32; ++ch) {
                        tmem_ld_wait();
                        if (ch + 1 < TC_BN / 32) tmem_ld32(taddr + (ch + 1) * 32, w[(ch + 1) & 1]);
                        uint32_t o = 0;
#pragma unroll
                        for (int e = 0; e < 32; ++e) o |= w[ch & 1][e];
                        if (o == 0x7fc12345u) cnt++;          // keeps the loads alive
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&t_empty[acc]);
                if (++acc == ACC) { acc = 0; accphase ^= 1; }
                continue;
            }
            uint32_t v[2][32];
            tmem_ld32(taddr, v[0]);
#pragma unroll
            for (int ch = 0; ch < TC_BN / 32; ++ch) {
                tmem_ld_wait();                                     // chunk ch is in v[ch & 1]
                if (ch + 1 < TC_BN / 32) tmem_ld32(taddr + (ch + 1) * 32, v[(ch + 1) & 1]);   // prefetch the next chunk
                const int jbase = kt * TC_BN + ch * 32;
                const float4 *rn = reinterpret_cast<const float4 *>(s_rn + (t % TC_NRING) * TC_BN + ch * 32);   // broadcast reads
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    const float4 n4 = rn[c4];
                    const float nn[4] = {n4.x, n4.y, n4.z, n4.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float d = fmaf(-2.f, __uint_as_float(v[ch & 1][c4 * 4 + e]), nn[e]);   // +inf for keys >= N
                        if (d < thr) {
                            buf[cnt] = make_uint2(__float_as_uint(d), (uint32_t)(jbase + c4 * 4 + e));
                            ++cnt;
                        }
                    }
                }
                // lists that could overflow during the next 32 columns are compacted now
                unsigned need = __ballot_sync(FULLW, cnt > TC_CAP - 32);
                while (need) {
                    const int r = __ffs(need) - 1;
                    need &= need - 1;
                    const int n_r = __shfl_sync(FULLW, cnt, r);
                    const float m_r = __shfl_sync(FULLW, margin, r);
                    const unsigned long long p_r = __shfl_sync(FULLW, (unsigned long long)buf, r);
                    __syncwarp();
                    int nc; float nt;
                    compact_row(reinterpret_cast<uint2 *>(p_r), n_r, a.k, m_r, lane, nc, nt);
                    if (lane == r) {
                        cnt = nc;
                        thr = nt;
                        if (nc > TC_CAP - 32) { ovf = true; cnt = 0; thr = -CUDART_INF_F; }
                    }
                }
            }
            // accumulator drained: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[acc]);
            if (++acc == ACC) { acc = 0; accphase ^= 1; }
        }

        // the last shrink of each list happens in the re-rank kernel (one warp per row, full occupancy)
        if (active) {
            if (a.split_counts) { a.cand_cnt[2 * grow] = ovf ? 0 : cnt; a.cand_cnt[2 * grow + 1] = 0; }
            else a.cand_cnt[grow] = ovf ? 0 : cnt;
            a.overflow[grow] = (ovf || cnt < a.k) ? 1 : 0;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, ACC * TC_BN);
    }
}

// ---------------------------------------------------------------------------------
// exact re-rank: one warp per query
// ---------------------------------------------------------------------------------
struct RerankArgs {
    const float *x_nc;     // [B][N][C]
    const float *norm;     // [B][N]
    const float *nmax;     // [B]
    const uint2 *cand;     // [B][N][TC_CAP]
    const int *cand_cnt;   // [B][N]
    const int *overflow;   // [B][N]
    int64_t *idx64;
    int32_t *idx32;
    int N, k, step, kout;
    int unordered;         // 1: the caller only needs the neighbour SET (EdgeConv is order-invariant)
    const int *perm;       // pruned path: rows and candidates are sorted positions, perm[b][s] = original index; else null
    int cap;               // entries per row in `cand`
    int split;             // 1: a row's list is two halves of cap / 2 entries with counts cand_cnt[2 row], cand_cnt[2 row + 1]
    int *big_list;         // [B][N] rows (scan order) with more than TC_CAP candidates, big_count [B] (zeroed by the host)
    int *big_count;
    int *fb_list;          // [B][N] rows left to the CUDA-core fallback (overflowed lists), fb_count [B] (zeroed by the host)
    int *fb_count;
};

template <int VEC>
__device__ __forceinline__ void load_row(const float *p, float (&v)[VEC]) {
    if constexpr (VEC == 2) {
        float2 t = __ldg(reinterpret_cast<const float2 *>(p)); v[0] = t.x; v[1] = t.y;
    } else {
        float4 t = __ldg(reinterpret_cast<const float4 *>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
}

// Bracket of the k-th smallest entry of a warp-distributed list (entries e = s*32 + lane, +inf padding, n > k of them
// finite): returns hi with count(d <= hi) >= k and lo with count(d <= lo) < k.  One pass: the entries are binned into 32
// equal-width buckets over [min, max] (a shared-memory histogram, one bucket per lane, prefix sums by shuffles); the
// bucket in which the running count reaches k holds the k-th entry, lo = the largest entry of the buckets before it,
// hi = the largest entry of the bucket itself -- both values of actual entries, so no rounding of the bucket edges can
// matter (the bucket index is a monotone function of the value).  The bracket holds the 2-3 entries of one bucket
// instead of exactly one, which the exact pass absorbs; the two-sided bisection this replaces took ~7 halvings of ~30
// instructions, a quarter of the kernel's instructions.
template <int SL>
__device__ __forceinline__ void kth_bracket(const float (&dv)[SL], int k, float &lo_out, float &hi_out, int *hist, int lane) {
    float mn = CUDART_INF_F, mx = -CUDART_INF_F;
#pragma unroll
    for (int s = 0; s < SL; ++s) {
        mn = fminf(mn, dv[s]);
        if (dv[s] < CUDART_INF_F) mx = fmaxf(mx, dv[s]);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(FULLW, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(FULLW, mx, o));
    }
    const float scale = mx > mn ? 32.f / (mx - mn) : 0.f;
    hist[lane] = 0;
    __syncwarp();
    int bk[SL];
#pragma unroll
    for (int s = 0; s < SL; ++s) {
        bk[s] = 32;                                        // padding: no bucket
        if (dv[s] < CUDART_INF_F) {
            bk[s] = min(31, max(0, (int)((dv[s] - mn) * scale)));
            atomicAdd(&hist[bk[s]], 1);
        }
    }
    __syncwarp();
    int pre = hist[lane];                                  // inclusive prefix count of bucket `lane`
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULLW, pre, o);
        if (lane >= o) pre += t;
    }
    const int B = __ffs(__ballot_sync(FULLW, pre >= k)) - 1;     // n > k entries are finite: some bucket reaches k
    float lo = -CUDART_INF_F, hi = -CUDART_INF_F;
#pragma unroll
    for (int s = 0; s < SL; ++s) {
        if (bk[s] < B) lo = fmaxf(lo, dv[s]);
        if (bk[s] <= B) hi = fmaxf(hi, dv[s]);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        lo = fmaxf(lo, __shfl_xor_sync(FULLW, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(FULLW, hi, o));
    }
    __syncwarp();
    lo_out = lo;                                           // -inf when the k-th entry lies in the first bucket
    hi_out = hi;
}

template <int C, int SL>
__device__ __forceinline__ void rerank_row(const RerankArgs &a, const uint2 *cand, int n, int n0, const int *perm, const float *xb,
                                           const float *nb, int q, float qn, const float (&qv)[16], float margin,
                                           size_t grow, unsigned long long *skey, int lane) {
    constexpr int LPC = C / 16;                         // lanes per candidate in the exact pass (16 channels each)
    constexpr int SCAP_ROW = SL * 32;
    int *sl = reinterpret_cast<int *>(skey);            // survivors' point ids, then (step 3) the keys overwrite both
    float *sd = reinterpret_cast<float *>(skey) + SCAP_ROW;
    // 1. last shrink of the list on the approximate distances (same bound + margin rule as the scan).
    //    In set-only mode entries with d~ <= lo - margin, where fewer than k entries have d~ <= lo, are
    //    certainly among the exact k nearest (anything that could beat them also lies below lo): they are
    //    written straight away and only the ambiguous band (lo - margin, hi + margin] is re-ranked.
    float ad[SL];
    int aj[SL];
#pragma unroll
    for (int s = 0; s < SL; ++s) {
        const int e = s * 32 + lane;
        ad[s] = CUDART_INF_F;
        aj[s] = 0;
        if (e < n) {
            uint2 t = cand[e < n0 ? e : e - n0 + (a.cap >> 1)];    // second half-list starts at cap / 2 (n0 = n: one list)
            ad[s] = __uint_as_float(t.x);
            aj[s] = (int)t.y;                                       // scan-order id; mapped through perm only if it survives
        }
    }
    float keep_below = CUDART_INF_F, sure_below = -CUDART_INF_F;
    if (n > a.k) {
        float lo_s, hi;
        kth_bracket<SL>(ad, a.k, lo_s, hi, reinterpret_cast<int *>(skey), lane);
        keep_below = hi + margin;
        if (a.unordered) sure_below = lo_s - margin;
    }
    int m = 0, n_sure = 0;
#pragma unroll
    for (int s = 0; s < SL; ++s) {
        const bool valid = ad[s] < CUDART_INF_F;
        const bool sure = valid && ad[s] <= sure_below;
        const bool keep = valid && !sure && ad[s] <= keep_below;
        const unsigned msk = __ballot_sync(FULLW, keep);
        const unsigned ssk = __ballot_sync(FULLW, sure);
        if (keep || sure) {
            const int jo = perm ? perm[aj[s]] : aj[s];              // original point index
            if (keep) sl[m + __popc(msk & ((1u << lane) - 1))] = jo;
            else {
                const size_t o = grow * a.kout + n_sure + __popc(ssk & ((1u << lane) - 1));
                if (a.idx64) a.idx64[o] = jo;
                if (a.idx32) a.idx32[o] = jo;
            }
        }
        m += __popc(msk);
        n_sure += __popc(ssk);
    }
    const int k_left = a.k - n_sure;                    // >= 1: fewer than k entries lie below lo
    __syncwarp();

    // 2. exact fp32 distances of the m survivors, 32 / LPC at a time: LPC lanes share a candidate, each with 16 channels
    //    as four 16-byte loads (the lanes of a candidate read LPC * 16 contiguous bytes per load), the query's matching
    //    channels stay in registers, log2(LPC) shuffles finish the dot product
    if constexpr (C == 3) {
        // xyz clouds: rows of (x, y, z, 0), one lane per candidate, the arithmetic of the brute-force scan
        // (knn_select.cu: fl(q0 r0), two fused steps, then fl(fl(|x_j|^2 - 2 t) + |x_i|^2))
        for (int e = lane; e < m; e += 32) {
            const int j = sl[e];
            const float4 r = __ldg(reinterpret_cast<const float4 *>(xb) + j);
            float t = __fmul_rn(qv[0], r.x);
            t = fmaf(qv[1], r.y, t);
            t = fmaf(qv[2], r.z, t);
            sd[e] = __fadd_rn(fmaf(-2.f, t, nb[j]), qn);
        }
    } else {
        constexpr int CPP = 32 / LPC;                   // candidates per pass
        const int u = lane / LPC, w = lane % LPC;
        for (int e0 = 0; e0 < m; e0 += CPP) {
            const int j = sl[min(e0 + u, m - 1)];
            const float4 *xr = reinterpret_cast<const float4 *>(xb + (size_t)j * C) + w;
            float4 xv[4];
#pragma unroll
            for (int it = 0; it < 4; ++it) xv[it] = __ldg(xr + it * LPC);
            float p0 = 0.f, p1 = 0.f;
#pragma unroll
            for (int it = 0; it < 4; it += 2) {
                p0 = fmaf(qv[it * 4 + 0], xv[it].x, p0);
                p1 = fmaf(qv[it * 4 + 4], xv[it + 1].x, p1);
                p0 = fmaf(qv[it * 4 + 1], xv[it].y, p0);
                p1 = fmaf(qv[it * 4 + 5], xv[it + 1].y, p1);
                p0 = fmaf(qv[it * 4 + 2], xv[it].z, p0);
                p1 = fmaf(qv[it * 4 + 6], xv[it + 1].z, p1);
                p0 = fmaf(qv[it * 4 + 3], xv[it].w, p0);
                p1 = fmaf(qv[it * 4 + 7], xv[it + 1].w, p1);
            }
            float kk = p0 + p1;
#pragma unroll
            for (int o = LPC / 2; o; o >>= 1) kk += __shfl_xor_sync(FULLW, kk, o);
            // reference arithmetic: fl(fl(|x_j|^2 - 2 t) + |x_i|^2)
            if (w == 0 && e0 + u < m) sd[e0 + u] = __fadd_rn(fmaf(-2.f, kk, nb[j]), qn);
        }
    }
    __syncwarp();

    // 3. rank the survivors by (distance, index) and write the k best in order.  (distance, index) is packed into one
    //    64-bit key -- the float mapped to an order-preserving unsigned -- so a comparison is one 64-bit compare and
    //    the broadcast read one 8-byte load
    unsigned long long key[SL];
    int rank[SL];
#pragma unroll
    for (int s = 0; s < SL; ++s) {
        const int e = s * 32 + lane;
        key[s] = ~0ull;
        rank[s] = 0;
        if (e < m) {
            unsigned bits = __float_as_uint(sd[e]);
            bits ^= (unsigned)((int)bits >> 31) | 0x80000000u;
            key[s] = ((unsigned long long)bits << 32) | (unsigned)sl[e];
        }
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < SL; ++s)
        if (s * 32 + lane < m) skey[s * 32 + lane] = key[s];
    __syncwarp();
    if (m <= 32) {
        for (int e = 0; e < m; ++e) rank[0] += (skey[e] < key[0]) ? 1 : 0;
    } else {
        for (int e = 0; e < m; ++e) {
            const unsigned long long ok = skey[e];
#pragma unroll
            for (int s = 0; s < SL; ++s) {
                if (s * 32 >= m) break;
                rank[s] += (ok < key[s]) ? 1 : 0;
            }
        }
    }
#pragma unroll
    for (int s = 0; s < SL; ++s) {
        const int e = s * 32 + lane;
        if (e < m && rank[s] < k_left && rank[s] % a.step == 0) {
            const size_t o = grow * a.kout + n_sure + rank[s] / a.step;
            const int jo = (int)(unsigned)key[s];
            if (a.idx64) a.idx64[o] = jo;
            if (a.idx32) a.idx32[o] = jo;
        }
    }
    (void)q;
}

// BIG = false: one warp per row of the scan; rows with more than TC_CAP candidates (dense near-tie regions on the
// pruned path, whose lists hold up to TCP_CAP) are only listed.  BIG = true: a second, usually empty launch takes
// the listed rows with the wide instantiation -- its registers and shared memory would otherwise cost every row
// an occupancy step.
template <int C, bool BIG>
__global__ void __launch_bounds__(256, BIG ? 1 : 6) knn_tc_rerank_kernel(RerankArgs a) {
    constexpr int LPC = C / 16;
    constexpr int XS = C == 3 ? 4 : C;                  // row stride of x_nc (xyz clouds: padded to 16 bytes)
    constexpr int SCAP = BIG ? TCP_CAP : TC_CAP;
    __shared__ unsigned long long s_key[8][SCAP];       // per warp: survivors' ids | exact distances, then the packed keys
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int *perm = a.perm ? a.perm + (size_t)b * a.N : nullptr;
    const float *xb = a.x_nc + (size_t)b * a.N * XS;
    const float *nb = a.norm + (size_t)b * a.N;
    const int rows = BIG ? a.big_count[b] : a.N;
    for (int slot = blockIdx.x * 8 + warp; slot < rows; slot += gridDim.x * 8) {
        const int qs = BIG ? a.big_list[(size_t)b * a.N + slot] : slot;       // row in the scan's order
        const size_t srow = (size_t)b * a.N + qs;
        // the row waits on chains of dependent loads: everything that only needs qs is issued first, everything that
        // needs the point id next, before any of it is tested
        const int q = perm ? perm[qs] : qs;               // original point index
        const int n0 = a.split ? a.cand_cnt[2 * srow] : a.cand_cnt[srow];
        const int n1 = a.split ? a.cand_cnt[2 * srow + 1] : 0;
        const float nmax = a.nmax[b];
        const size_t grow = (size_t)b * a.N + q;
        const int ovf = BIG ? 0 : a.overflow[grow];
        float qv[16];                                    // this lane's 16 channels of the query (layout of the exact pass)
        if constexpr (C == 3) {
            const float4 t = __ldg(reinterpret_cast<const float4 *>(xb) + q);
            qv[0] = t.x; qv[1] = t.y; qv[2] = t.z;
        } else {
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const float4 t = __ldg(reinterpret_cast<const float4 *>(xb + (size_t)q * C) + it * LPC + lane % LPC);
                qv[it * 4] = t.x; qv[it * 4 + 1] = t.y; qv[it * 4 + 2] = t.z; qv[it * 4 + 3] = t.w;
            }
        }
        const float qn = nb[q];
        if (ovf) {                                        // the fallback kernel writes this row
            if (lane == 0) a.fb_list[(size_t)b * a.N + atomicAdd(&a.fb_count[b], 1)] = q;
            continue;
        }
        const int n = n0 + n1;
        if (!BIG && n > TC_CAP) {
            if (lane == 0) a.big_list[(size_t)b * a.N + atomicAdd(&a.big_count[b], 1)] = qs;
            continue;
        }
        const uint2 *cand = a.cand + srow * a.cap;
        const float margin = TC_MARGIN * sqrtf(qn * nmax);
        if constexpr (BIG) {
            rerank_row<C, TCP_CAP / 32>(a, cand, n, n0, perm, xb, nb, q, qn, qv, margin, grow, s_key[warp], lane);
        } else {
            // short lists (the pruned scan's fixed thresholds leave ~1.5 k entries) take the narrow instantiation
            if (n <= 96) rerank_row<C, 3>(a, cand, n, n0, perm, xb, nb, q, qn, qv, margin, grow, s_key[warp], lane);
            else if (n <= 128) rerank_row<C, 4>(a, cand, n, n0, perm, xb, nb, q, qn, qv, margin, grow, s_key[warp], lane);
            else rerank_row<C, 8>(a, cand, n, n0, perm, xb, nb, q, qn, qv, margin, grow, s_key[warp], lane);
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------
// pruned scan (default for N >= TCP_MIN_N, k <= 64): the cloud is sorted along a Morton curve in the
// space of its three leading principal directions, 64-key tiles get an AABB in that space, and a
// query tile only visits key tiles whose lower bound can still beat its largest threshold.
//
// Bound: for orthonormal directions v_1..v_3, |V^T(q - x)|^2 <= |q - x|^2, hence
//   |q - x|^2 >= dist^2(AABB(V^T q : q in query tile), AABB(V^T x : x in key tile)) =: LB.
// The directions only have to be orthonormal (checked in the PCA kernel, else they are zeroed =
// no pruning); how well they capture the variance only decides how much is skipped.
// ---------------------------------------------------------------------------------
constexpr int TCP_MIN_N = 1024;
constexpr int TCP_FULL_MIN_N = 32768;  // unstructured clouds at least this large take the single-pass full scan
constexpr int TCP_SPLIT = 8;          // partial Gram matrices per cloud
constexpr int TCP_PRE = 8;            // nearest tiles every query tile reads unconditionally (first bound refresh after them)
constexpr int TCP_ITERS = 5;          // subspace iterations (lambda_3 / lambda_4 is ~4 on layer activations: 4^5 = 1000x)
constexpr uint32_t TCP_END = 0xffffffffu;
constexpr int TCP_THREADS = 320;      // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quarter)
constexpr int TCP_HCAP = TCP_CAP / 2; // each of a row's two epilogue threads owns half of its candidate list
// TMEM accumulator stages (64 columns each) behind the query tile, which lives in TMEM as well (C columns: hi | lo):
// C = 64: 64 + 3 x 64 = 256 columns, so two CTAs still share an SM's 512; C = 128: 128 + 4 x 64 = 384 (one CTA per SM)
// C = 3 (xyz clouds): the query tile is one K = 16 step = 8 columns (32 reserved): 32 + 3 x 64 = 224 of 256
__host__ __device__ constexpr int tcp_acc(int C) { return C == 128 ? 4 : 3; }
__host__ __device__ constexpr int tcp_tmem_cols(int C) { return C == 128 ? 512 : 256; }
constexpr int TCP_NRING = 16;         // key-norm ring: the producer runs at most STAGES + ACC + 1 tiles ahead of the epilogue
__host__ __device__ constexpr int tcp_stages(int C) { return C == 128 ? 3 : 4; }

// part[b][sp][C*C + C] = (sum_n x x^T | sum_n x) over the sp-th slice of a strided subsample of the cloud's points
// (every `stride`-th point, Ns of them: the directions only steer the pruning, a sample is enough)
template <int C>
__global__ void __launch_bounds__(256) tcp_moments_kernel(const float *__restrict__ x, float *__restrict__ part, int N,
                                                          int Ns, int stride) {
    constexpr int R = C / 16;
    __shared__ float t[C][33];
    const int b = blockIdx.y, sp = blockIdx.x, tid = threadIdx.x;
    const int chunk = (Ns + TCP_SPLIT - 1) / TCP_SPLIT;
    const int n_lo = sp * chunk, n_hi = min(Ns, n_lo + chunk);
    const float *xb = x + (size_t)b * C * N;
    const int ti = tid >> 4, tj = tid & 15;
    float acc[R][R];
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < R; ++j) acc[i][j] = 0.f;
    float sum = 0.f;
    for (int n0 = n_lo; n0 < n_hi; n0 += 32) {
        for (int c = tid >> 5; c < C; c += 8) {
            const int n = n0 + (tid & 31);
            t[c][tid & 31] = n < n_hi ? xb[(size_t)c * N + (size_t)n * stride] : 0.f;
        }
        __syncthreads();
#pragma unroll 4
        for (int p = 0; p < 32; ++p) {
            float av[R], bv[R];
#pragma unroll
            for (int i = 0; i < R; ++i) { av[i] = t[ti * R + i][p]; bv[i] = t[tj * R + i][p]; }
#pragma unroll
            for (int i = 0; i < R; ++i)
#pragma unroll
                for (int j = 0; j < R; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (tid < C) {
#pragma unroll 8
            for (int p = 0; p < 32; ++p) sum += t[tid][p];
        }
        __syncthreads();
    }
    float *o = part + ((size_t)b * TCP_SPLIT + sp) * (C * C + C);
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < R; ++j) o[(ti * R + i) * C + tj * R + j] = acc[i][j];
    if (tid < C) o[C * C + tid] = sum;
}

// pca[b] = { V[3][C], m[3] = V mu, sigma_1 }  (3*C + 4 floats): three orthonormal directions spanning (approximately)
// the leading principal subspace of the cloud, by subspace iteration on its covariance.  One CTA per cloud.
// structured[b] = 1: the box-pruned two-pass scan takes the cloud.  Large clouds (n_full >= TCP_FULL_MIN_N) whose three
// directions carry less than half of the variance (e.g. i.i.d. Gaussian features: nothing can be skipped, and at
// that size the repeated tensor-core pass costs more than the list compactions it avoids) get 0: their directions
// are zeroed -- all sort keys equal, the stable sort keeps the original order -- and the single-pass full scan
// takes them.  (At 10 000 points the two-pass scan is the faster one even without any structure: 2.3 vs 3.0 ms.)
template <int C>
__global__ void __launch_bounds__(256) tcp_pca_kernel(const float *__restrict__ part, float *__restrict__ pca,
                                                      int *__restrict__ structured, int N, int n_full) {
    extern __shared__ float sm[];
    float *cov = sm;                  // [C][C+1]
    float *mu = cov + C * (C + 1);    // [C]
    float *V = mu + C;                // [3][C]
    float *W = V + 3 * C;             // [3][C]
    __shared__ float s_lambda[3];
    __shared__ int s_ok;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const float *pb = part + (size_t)b * TCP_SPLIT * (C * C + C);
    const float invN = 1.f / (float)N;
    if (tid < C) {
        float s = 0.f;
        for (int sp = 0; sp < TCP_SPLIT; ++sp) s += pb[(size_t)sp * (C * C + C) + C * C + tid];
        mu[tid] = s * invN;
    }
    __syncthreads();
    for (int e = tid; e < C * C; e += 256) {
        float s = 0.f;
        for (int sp = 0; sp < TCP_SPLIT; ++sp) s += pb[(size_t)sp * (C * C + C) + e];
        const int r = e / C, c = e % C;
        cov[r * (C + 1) + c] = s * invN - mu[r] * mu[c];
    }
    for (int e = tid; e < 3 * C; e += 256) {
        unsigned h = (unsigned)e * 2654435761u + 12345u;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        W[e] = (float)(h & 0xffff) * (2.f / 65535.f) - 1.f;
    }
    __syncthreads();

    // Gram-Schmidt of W into V by warp 0 (lane holds C/32 entries of each vector)
    auto orthonormalize = [&]() {
        if (tid < 32) {
            constexpr int E = C / 32;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                float w[E];
#pragma unroll
                for (int e = 0; e < E; ++e) w[e] = W[i * C + lane + 32 * e];
                for (int pass = 0; pass < 2; ++pass) {        // twice: re-orthogonalisation keeps |V^T V - I| at fp32 level
                    for (int j = 0; j < i; ++j) {
                        float d = 0.f;
#pragma unroll
                        for (int e = 0; e < E; ++e) d = fmaf(w[e], V[j * C + lane + 32 * e], d);
                        for (int o = 16; o; o >>= 1) d += __shfl_xor_sync(FULLW, d, o);
#pragma unroll
                        for (int e = 0; e < E; ++e) w[e] = fmaf(-d, V[j * C + lane + 32 * e], w[e]);
                    }
                }
                float nn = 0.f;
#pragma unroll
                for (int e = 0; e < E; ++e) nn = fmaf(w[e], w[e], nn);
                for (int o = 16; o; o >>= 1) nn += __shfl_xor_sync(FULLW, nn, o);
                const float inv = nn > 1e-30f ? rsqrtf(nn) : 0.f;
                if (lane == 0) s_lambda[i] = sqrtf(nn);
#pragma unroll
                for (int e = 0; e < E; ++e) V[i * C + lane + 32 * e] = w[e] * inv;
                __syncwarp();
            }
        }
        __syncthreads();
    };
    orthonormalize();
    for (int it = 0; it < TCP_ITERS; ++it) {
        for (int e = tid; e < 3 * C; e += 256) {
            const int i = e / C, r = e % C;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 4
            for (int c = 0; c < C; c += 4) {
                s0 = fmaf(cov[r * (C + 1) + c], V[i * C + c], s0);
                s1 = fmaf(cov[r * (C + 1) + c + 1], V[i * C + c + 1], s1);
                s2 = fmaf(cov[r * (C + 1) + c + 2], V[i * C + c + 2], s2);
                s3 = fmaf(cov[r * (C + 1) + c + 3], V[i * C + c + 3], s3);
            }
            W[e] = (s0 + s1) + (s2 + s3);
        }
        __syncthreads();
        orthonormalize();
    }
    // validity: V must be orthonormal for the bound to hold; anything else (degenerate cloud, NaN input) -> no pruning
    if (tid < 32) {
        bool ok = true;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j <= i; ++j) {
                float d = 0.f;
                for (int c = lane; c < C; c += 32) d = fmaf(V[i * C + c], V[j * C + c], d);
                for (int o = 16; o; o >>= 1) d += __shfl_xor_sync(FULLW, d, o);
                const float want = i == j ? 1.f : 0.f;
                if (!(fabsf(d - want) <= 1e-4f)) ok = false;
            }
        // explained variance: lambda_i ~ |cov v_i| of the last iteration against the trace
        float tr = 0.f;
        for (int c = lane; c < C; c += 32) tr += cov[c * (C + 1) + c];
        for (int o = 16; o; o >>= 1) tr += __shfl_xor_sync(FULLW, tr, o);
        if (n_full >= TCP_FULL_MIN_N && !(s_lambda[0] + s_lambda[1] + s_lambda[2] >= 0.5f * tr)) ok = false;
        if (lane == 0) { s_ok = ok ? 1 : 0; structured[blockIdx.x] = ok ? 1 : 0; }
    }
    __syncthreads();
    float *o = pca + (size_t)b * (3 * C + 4);
    for (int e = tid; e < 3 * C; e += 256) o[e] = s_ok ? V[e] : 0.f;
    if (tid < 3) {
        float m = 0.f;
        for (int c = 0; c < C; ++c) m = fmaf(V[tid * C + c], mu[c], m);
        o[3 * C + tid] = s_ok ? m : 0.f;
    }
    if (tid == 3) o[3 * C + 3] = s_ok ? fmaxf(s_lambda[0], fmaxf(s_lambda[1], s_lambda[2])) : 0.f;   // = sigma_1^2 (largest |cov v|)
}

// proj[i][b][n] = v_i . x_n ; sort key = (cloud, 30-bit Morton code of the projections in a +-4 sigma_1 cube)
template <int C>
__global__ void __launch_bounds__(128) tcp_project_kernel(const float *__restrict__ x, const float *__restrict__ pca,
                                                          float *__restrict__ proj, unsigned *__restrict__ keys,
                                                          int *__restrict__ vals, float *__restrict__ norm, int B, int N, int bits) {
    __shared__ float V[3 * C + 4];
    const int b = blockIdx.y;
    for (int e = threadIdx.x; e < 3 * C + 4; e += 128) V[e] = pca[(size_t)b * (3 * C + 4) + e];
    __syncthreads();
    const int n = blockIdx.x * 128 + threadIdx.x;
    if (n >= N) return;
    const float *xb = x + (size_t)b * C * N + n;
    float p0 = 0.f, p1 = 0.f, p2 = 0.f, sq = 0.f;
#pragma unroll 8
    for (int c = 0; c < C; ++c) {
        const float v = xb[(size_t)c * N];
        p0 = fmaf(V[c], v, p0);
        p1 = fmaf(V[C + c], v, p1);
        p2 = fmaf(V[2 * C + c], v, p2);
        sq = __fadd_rn(sq, __fmul_rn(v, v));                    // |x|^2 in sqnorm_kernel's order (the point is read here anyway)
    }
    const size_t o = (size_t)b * N + n;
    norm[o] = sq;
    proj[o] = p0;
    proj[(size_t)B * N + o] = p1;
    proj[2 * (size_t)B * N + o] = p2;
    const float sig = sqrtf(fmaxf(V[3 * C + 3], 0.f));
    const float sc = sig > 0.f ? 1024.f / (8.f * sig) : 0.f;
    const float pv[3] = {p0, p1, p2};
    unsigned cell[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float t = (pv[i] - V[3 * C + i]) * sc + 512.f;
        t = t == t ? t : 0.f;                                   // NaN -> 0
        cell[i] = ((unsigned)fminf(fmaxf(t, 0.f), 1023.f)) >> (10 - bits);
    }
    const unsigned code = hilbert3(cell[0], cell[1], cell[2], bits);
    keys[o] = ((unsigned)b << (3 * bits)) | code;              // cloud | Hilbert-curve position with `bits` bits per axis
    vals[o] = n;
}

// Extension operand of a 64-key tile: [64 keys][16 bf16] in the interleaved (no-swizzle) K-major layout of
// make_kmajor_interleaved_desc -- 8-row groups of 256 bytes, inside a group the two 16-byte column blocks 128 bytes apart.
constexpr int TCP_KEXT_BYTES = TC_BN * 32;
__host__ __device__ __forceinline__ uint32_t tcp_kext_offset(int r, int chunk) {
    return (uint32_t)((r >> 3) * 256 + chunk * 128 + (r & 7) * 16);
}

// one warp per 64-key tile of the sorted order: perm / inverse permutation, sorted key norms (+inf padding),
// tile AABB in the projected space
//
// XYZ (C = 3 clouds, the coordinates themselves are the "projections"): `proj` is the cloud x[B][C][N]; the K = 16 operand
// row of a key holds the whole bf16 x 3 product AND the norm -- (hi, lo, hi, |k|^2 hi, mid, lo, 0 ...) against the query
// row -2 (hi, hi, lo), 1, 1, 1 -- so ONE MMA per key tile yields |k|^2 - 2 q.k; x4[b][o] = (x, y, z, 0) in the original
// order feeds the exact re-rank, and every cloud is flagged structured.
struct TcpTileArgs {
    const int *sorted_vals;   // [B][N] point of every sorted position
    const float *norm;        // [B][N]
    const float *proj;        // projections [3][B][N], or (XYZ) the cloud itself [B][C][N]
    int *perm, *inv;
    float *norm_pad, *boxes, *boxes32;
    unsigned *nmax_bits;
    uint8_t *kext;
    int B, N, Npad, tiles, C;
    float4 *x4;               // XYZ: [B][N] (x, y, z, 0) in the original order
    int *structured;          // XYZ: every cloud is flagged
};

// one warp, one 64-key tile t of cloud b
template <bool XYZ>
__device__ __forceinline__ void tcp_tile_work(const TcpTileArgs &a, int b, int t, int lane) {
    const int *sorted_vals = a.sorted_vals;
    const float *norm = a.norm, *proj = a.proj;
    int *perm = a.perm, *inv = a.inv;
    float *norm_pad = a.norm_pad, *boxes = a.boxes, *boxes32 = a.boxes32;
    unsigned *nmax_bits = a.nmax_bits;
    uint8_t *kext = a.kext;
    const int B = a.B, N = a.N, Npad = a.Npad, tiles = a.tiles, C = a.C;
    float4 *x4 = a.x4;
    float mn[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, mx[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    float nm = 0.f;
    {
#pragma unroll
    for (int h = 0; h < TC_BN / 32; ++h) {
        const int s = t * TC_BN + h * 32 + lane;
        float hmn[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, hmx[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
        float nkey = CUDART_INF_F;                             // +inf masks the zero-filled keys past N
        if (s < N) {
            const int o = sorted_vals[(size_t)b * N + s];
            perm[(size_t)b * N + s] = o;
            inv[(size_t)b * N + o] = s;
            const float nv = norm[(size_t)b * N + o];
            norm_pad[(size_t)b * Npad + s] = nv;
            nkey = nv;
            nm = fmaxf(nm, nv);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const float p = XYZ ? proj[((size_t)b * C + i) * N + o] : proj[(size_t)i * B * N + (size_t)b * N + o];
                hmn[i] = p;
                hmx[i] = p;
            }
            if (XYZ) x4[(size_t)b * N + o] = make_float4(hmn[0], hmn[1], hmn[2], 0.f);
        } else if (s < Npad) {
            norm_pad[(size_t)b * Npad + s] = CUDART_INF_F;
        }
        {
            // |k|^2 as three bf16 terms (hi + mid + lo = all 24 bits of the fp32 norm) in column 0..2 of a K = 16 operand
            // step: the scan adds the norm to -2 q.k on the tensor cores (one more MMA per tile).  Written as the
            // shared-memory image of the tile (interleaved K-major, tcp_kext_offset), fetched with one bulk copy.
            const int r = h * 32 + lane;
            __nv_bfloat16 e0 = __float2bfloat16_rn(nkey), e1 = __float2bfloat16_rn(0.f), e2 = e1;
            if (nkey < CUDART_INF_F) {
                const float r1 = nkey - __bfloat162float(e0);
                e1 = __float2bfloat16_rn(r1);
                e2 = __float2bfloat16_rn(r1 - __bfloat162float(e1));
            }
            uint4 c0, c1 = make_uint4(0u, 0u, 0u, 0u);
            if (XYZ) {
                // columns 0-2 hi, 3-5 lo, 6-8 hi, 9-11 the norm; keys past N: zero coordinates (hmn is +inf there), norm +inf
                unsigned h[3], l[3];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const float v = s < N ? hmn[i] : 0.f;        // hmn still holds this point's own coordinates here
                    const __nv_bfloat16 vh = __float2bfloat16_rn(v);
                    h[i] = __bfloat16_as_ushort(vh);
                    l[i] = __bfloat16_as_ushort(__float2bfloat16_rn(v - __bfloat162float(vh)));
                }
                c0.x = h[0] | (h[1] << 16);
                c0.y = h[2] | (l[0] << 16);
                c0.z = l[1] | (l[2] << 16);
                c0.w = h[0] | (h[1] << 16);
                c1.x = h[2] | ((unsigned)__bfloat16_as_ushort(e0) << 16);
                c1.y = (unsigned)__bfloat16_as_ushort(e1) | ((unsigned)__bfloat16_as_ushort(e2) << 16);
            } else {
                c0.x = (unsigned)__bfloat16_as_ushort(e0) | ((unsigned)__bfloat16_as_ushort(e1) << 16);
                c0.y = (unsigned)__bfloat16_as_ushort(e2);
                c0.z = 0u; c0.w = 0u;
            }
            uint8_t *img = kext + ((size_t)b * tiles + t) * TCP_KEXT_BYTES;
            *reinterpret_cast<uint4 *>(img + tcp_kext_offset(r, 0)) = c0;
            *reinterpret_cast<uint4 *>(img + tcp_kext_offset(r, 1)) = c1;
        }
        // box of this 32-point half (the rows one epilogue warp of the scan owns), then merged into the tile's box
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            for (int o = 16; o; o >>= 1) {
                hmn[i] = fminf(hmn[i], __shfl_xor_sync(FULLW, hmn[i], o));
                hmx[i] = fmaxf(hmx[i], __shfl_xor_sync(FULLW, hmx[i], o));
            }
            mn[i] = fminf(mn[i], hmn[i]);
            mx[i] = fmaxf(mx[i], hmx[i]);
        }
        if (lane == 0) {
            float *hb = boxes32 + ((size_t)b * tiles * 2 + t * 2 + h) * 6;
            hb[0] = hmn[0]; hb[1] = hmn[1]; hb[2] = hmn[2]; hb[3] = hmx[0]; hb[4] = hmx[1]; hb[5] = hmx[2];
        }
    }
    for (int o = 16; o; o >>= 1) nm = fmaxf(nm, __shfl_xor_sync(FULLW, nm, o));
    if (lane == 0) {
        atomicMax(nmax_bits + b, __float_as_uint(nm));       // norms are >= 0: the bit patterns order like the values
        float *bx = boxes + ((size_t)b * tiles + t) * 6;
        bx[0] = mn[0]; bx[1] = mn[1]; bx[2] = mn[2]; bx[3] = mx[0]; bx[4] = mx[1]; bx[5] = mx[2];
    }
    }
}

template <bool XYZ>
__global__ void __launch_bounds__(256) tcp_tiles_kernel(TcpTileArgs a) {
    const int b = blockIdx.y;
    if (XYZ && blockIdx.x == 0 && threadIdx.x == 0) a.structured[b] = 1;
    const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (t < a.tiles) tcp_tile_work<XYZ>(a, b, t, threadIdx.x & 31);
}

// work[rank] = i for query tile i = b * qtiles + qt, ranked by the squared diagonal of its bounding box in the
// projected space (wkey, written by tcp_tiles_kernel), largest first.  Every thread ranks one tile against all.
constexpr int TCP_MAX_WORK = 65536;
// wkey[b * qtiles + qt] = estimate of the work of a query tile: the number of key tiles whose lower bound lies within
// a quarter of the squared diagonal of the query tile's own box (rank correlation with the tiles it ends up visiting:
// 0.96 on layer activations; the diagonal alone: 0.7-0.8), ties by the diagonal.  One warp per query tile.
__device__ __forceinline__ void tcp_work_key_one(const float *__restrict__ boxes, unsigned *__restrict__ wkey, int tiles, int qtiles,
                                                 int b, int qt, int lane) {
    const float *bx = boxes + (size_t)b * tiles * 6;
    const int t0 = 2 * qt, t1 = min(t0 + 1, tiles - 1);
    float qlo[3], qhi[3], d2 = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        qlo[i] = fminf(bx[t0 * 6 + i], bx[t1 * 6 + i]);
        qhi[i] = fmaxf(bx[t0 * 6 + 3 + i], bx[t1 * 6 + 3 + i]);
        d2 = fmaf(qhi[i] - qlo[i], qhi[i] - qlo[i], d2);
    }
    if (!(d2 >= 0.f)) d2 = 0.f;
    int cnt = 0;
    for (int t = lane; t < tiles; t += 32) {
        float lb = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float g = fmaxf(fmaxf(qlo[i] - bx[t * 6 + 3 + i], bx[t * 6 + i] - qhi[i]), 0.f);
            lb = fmaf(g, g, lb);
        }
        cnt += (lb <= 0.25f * d2) ? 1 : 0;
    }
    cnt = __reduce_add_sync(FULLW, cnt);
    if (lane == 0) wkey[(size_t)b * qtiles + qt] = ((unsigned)min(cnt, 65535) << 16) | (__float_as_uint(d2) >> 16);
}
__global__ void __launch_bounds__(256) tcp_work_key_kernel(const float *__restrict__ boxes, unsigned *__restrict__ wkey,
                                                           int tiles, int qtiles) {
    const int qt = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (qt < qtiles) tcp_work_key_one(boxes, wkey, tiles, qtiles, blockIdx.y, qt, threadIdx.x & 31);
}

__global__ void __launch_bounds__(256) tcp_work_order_kernel(const unsigned *__restrict__ wkey, int *__restrict__ work, int n) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;    // one warp per tile
    if (i >= n) return;
    const unsigned mine = wkey[i];
    int rank = 0;
    for (int j = lane; j < n; j += 32) {
        const unsigned o = __ldg(wkey + j);
        rank += (o > mine || (o == mine && j < i)) ? 1 : 0;
    }
    rank = __reduce_add_sync(FULLW, rank);
    if (lane == 0) work[rank] = i;
}

// Clouds of at most 10 240 points: ONE CTA sorts a cloud's (code, point) pairs in shared memory -- five 4-bit passes of
// cub::BlockRadixSort over the 18 code bits instead of the device-wide sort's histogram + scan + three onesweep passes
// over all clouds (25 us against 45: those are latency-bound at 160 000 keys) -- and does the small per-cloud chores that
// used to be launches of their own on the way: zeroing the per-cloud counters and, for xyz clouds, the squared norms, the
// bounding box and the curve positions themselves.  Both sorts are stable LSD sorts of the same keys, so the permutation
// is the same; padding keys carry all-ones code bits and, being last in the input, stay behind every real key.
constexpr int TCP_SORT_ITEMS = 10;
constexpr int TCP_SORT_MAX_N = 1024 * TCP_SORT_ITEMS;
struct TcpCloudPrepArgs {
    TcpTileArgs t;
    const unsigned *keys;     // feature clouds: (cloud << code_bits) | code from tcp_project_kernel; xyz clouds: unused
    float *norm_out;          // xyz clouds: [B][N] squared norms (reference order), written here
    int *vals_out;            // [B][N] = t.sorted_vals
    int *fb_count, *big_count;  // [B] each, zeroed here
    int code_bits, axis_bits;
};
template <bool XYZ>
__global__ void __launch_bounds__(1024) tcp_cloud_prep_kernel(TcpCloudPrepArgs a) {
    using Sort = cub::BlockRadixSort<unsigned, 1024, TCP_SORT_ITEMS, int>;
    __shared__ typename Sort::TempStorage temp;
    __shared__ float s_red[6][32];
    __shared__ float s_bb[6];
    const int b = blockIdx.x, N = a.t.N;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        a.t.nmax_bits[b] = 0u;
        a.fb_count[b] = 0;
        a.big_count[b] = 0;
        if (XYZ) a.t.structured[b] = 1;
    }
    unsigned k[TCP_SORT_ITEMS];
    int v[TCP_SORT_ITEMS];
    if constexpr (XYZ) {
        // squared norms (sqnorm_kernel's order), bounding box, then the Hilbert-curve position inside it (xyz_code_kernel)
        const float *xb = a.t.proj + (size_t)b * a.t.C * N;
        float mn[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, mx[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
#pragma unroll
        for (int i = 0; i < TCP_SORT_ITEMS; ++i) {
            const int n = threadIdx.x * TCP_SORT_ITEMS + i;
            if (n < N) {
                float sq = 0.f;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float x = xb[(size_t)c * N + n];
                    sq = __fadd_rn(sq, __fmul_rn(x, x));
                    mn[c] = fminf(mn[c], x);
                    mx[c] = fmaxf(mx[c], x);
                }
                a.norm_out[(size_t)b * N + n] = sq;
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            for (int o = 16; o; o >>= 1) {
                mn[c] = fminf(mn[c], __shfl_xor_sync(FULLW, mn[c], o));
                mx[c] = fmaxf(mx[c], __shfl_xor_sync(FULLW, mx[c], o));
            }
            if (lane == 0) { s_red[c][warp] = mn[c]; s_red[3 + c][warp] = mx[c]; }
        }
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                float r = s_red[c][lane];
                for (int o = 16; o; o >>= 1) {
                    const float w = __shfl_xor_sync(FULLW, r, o);
                    r = c < 3 ? fminf(r, w) : fmaxf(r, w);
                }
                if (lane == 0) s_bb[c] = r;
            }
        }
        __syncthreads();
        const float ext = fmaxf(fmaxf(s_bb[3] - s_bb[0], s_bb[4] - s_bb[1]), fmaxf(s_bb[5] - s_bb[2], 1e-30f));
        const float top = (float)((1 << a.axis_bits) - 1);
        const float sc = top / ext;
#pragma unroll
        for (int i = 0; i < TCP_SORT_ITEMS; ++i) {
            const int n = threadIdx.x * TCP_SORT_ITEMS + i;
            k[i] = 0xffffffffu;
            v[i] = n;
            if (n < N) {
                unsigned q[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float t = (xb[(size_t)c * N + n] - s_bb[c]) * sc;
                    q[c] = (unsigned)fminf(fmaxf(t, 0.f), top);
                }
                k[i] = hilbert3(q[0], q[1], q[2], a.axis_bits);
            }
        }
    } else {
        const unsigned mask = a.code_bits >= 32 ? 0xffffffffu : ((1u << a.code_bits) - 1u);
#pragma unroll
        for (int i = 0; i < TCP_SORT_ITEMS; ++i) {
            const int n = threadIdx.x * TCP_SORT_ITEMS + i;                 // blocked arrangement = input order
            k[i] = n < N ? (a.keys[(size_t)b * N + n] & mask) : 0xffffffffu;
            v[i] = n;
        }
    }
    Sort(temp).Sort(k, v, 0, a.code_bits);
#pragma unroll
    for (int i = 0; i < TCP_SORT_ITEMS; ++i) {
        const int p = threadIdx.x * TCP_SORT_ITEMS + i;
        if (p < N) a.vals_out[(size_t)b * N + p] = v[i];
    }
    // (the tile pass stays a kernel of its own: 157 tiles of dependent gathers on ONE CTA per cloud took 60 us, spread over
    // the GPU they take 9)
}

struct TcpScanArgs {
    const float *norm_pad;  // [B][Npad] key norms in sorted order, +inf past N
    int Npad;
    const __nv_bfloat16 *xs; // [B][N][2C] hi | lo operands in sorted order (the array the key tensor map covers)
    const uint8_t *kext;    // [B][tiles][TCP_KEXT_BYTES] the same norms as a K = 16 bf16 operand step (tcp_tiles_kernel)
    const float *nmax;      // [B]
    const float *boxes;     // [B][tiles][6]
    const float *boxes32;   // [B][2 * tiles][6] boxes of the 32-point halves (= the rows of one epilogue warp)
    const int *perm;        // [B][N] sorted position -> original index
    uint2 *cand;            // [B][N][TCP_CAP] rows in sorted order, key ids are sorted positions
    int *cand_cnt;          // [B][N]          rows in sorted order
    int *overflow;          // [B][N]          rows in ORIGINAL order (consumed by the fallback scan)
    int *visited;           // [B][query tiles] statistics: key tiles scanned in the main pass (may be null)
    const int *work;        // [B * qtiles] query tiles (b * qtiles + qt) in launch order, or null = natural order
    const int *structured;  // [B] 1 = this kernel takes the cloud, 0 = the full scan does
    int N, k, tiles, pre, P;  // P = tiles rounded up to a power of two (sort width)
    int qtiles;
    long long *prof;        // [B * qtiles][16] per-CTA cycle counters (measurement builds, GCANET_TC_PROF=1), else null
    int refresh0, refresh_mul, refresh_max;   // pass A publishes its bound after refresh0 tiles, then every refresh_mul times as many, up to refresh_max
};

#ifdef GCANET_MEASUREMENT_AIDS
#define TCP_PROF_T0() const long long prof_t0__ = clock64()
#define TCP_PROF_ADD(acc) acc += clock64() - prof_t0__
#define TCP_PROF_NOW() clock64()
#else
#define TCP_PROF_T0() do { } while (0)
#define TCP_PROF_ADD(acc) do { } while (0)
#define TCP_PROF_NOW() 0ll
#endif

// SM = 1: k <= 64, one minimum per column slot.  SM = 2: k <= 128, the two smallest per slot (128 distinct keys; the
// 64 extra registers cost the second resident CTA).
//
// C = 3 (XYZ): the cloud's own coordinates take the place of the principal directions, and the whole distance is ONE
// K = 16 MMA per key tile -- the 2 KB operand image written by tcp_tiles_kernel<true> holds the bf16 x 3 split of the
// three coordinates and the norm; no tensor map, no swizzled operand blocks, the query row is built from the same image.
template <int C, int SM>
__global__ void __launch_bounds__(TCP_THREADS, C == 128 ? 1 : 2)
knn_tcp_scan_kernel(const __grid_constant__ CUtensorMap tmap_k, TcpScanArgs a) {
    constexpr bool XYZ = (C == 3);
    constexpr int NBLK = XYZ ? 0 : 2 * C / TC_KB;
    constexpr int NH = XYZ ? 0 : C / TC_KB;
    constexpr int BLK_BYTES = TC_BN * 128;
    constexpr int TILE_BYTES = NBLK * BLK_BYTES;
    // The hand-offs (TMA -> MMA -> epilogue -> MMA -> TMA) each cost a barrier round trip of ~1 us under load, so the
    // rings are deep: the epilogue should never wait for an accumulator.  C = 64: 32 + 4*16 KB -> two CTAs per SM.
    constexpr int STAGES = tcp_stages(C);
    constexpr int ACC = tcp_acc(C);
    constexpr int ACOLS = XYZ ? 32 : C;                     // TMEM columns of the query tile: C/2 hi, C/2 lo (two bf16 each)

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *sB = smem;
    // The query tile never touches shared memory: each row is read once from global memory, scaled by -2 (exact in
    // bf16) and stored into TMEM, from where every MMA takes its A operand.  With both operands in shared memory an
    // M = 128, N = 64 step reads 6 KB per 32-cycle tensor slot -- 1.5x what shared memory delivers (measured: the MMA
    // issuer busy 85 % of a CTA's life, tensor pipe 45 % active); with A in TMEM it reads 2 KB.
    // The key norms ride on the tensor cores too: one extra K = 16 step multiplies a constant (1, 1, 1, 0, ...) query
    // row with (|k|^2 hi, mid, lo, 0, ...) of every key, so the accumulator holds  |k|^2 - 2 q.k  itself.
    uint8_t *sKe = sB + STAGES * TILE_BYTES;                // [STAGES][TCP_KEXT_BYTES] norm operand of the stage's key tile
    uint8_t *sQe = sKe + STAGES * TCP_KEXT_BYTES;           // 256 B: one 8-row group of the constant query operand (SBO = 0)
    uint64_t *bars = reinterpret_cast<uint64_t *>(sQe + 256);
    uint64_t *full = bars;
    uint64_t *empty = bars + STAGES;
    uint64_t *t_full = bars + 2 * STAGES;
    uint64_t *t_empty = t_full + ACC;
    uint64_t *thr_ready = t_empty + ACC;                    // [1] epilogue -> producer: final thresholds of pass A are published
    uint64_t *a_ready = thr_ready + 1;                      // [1] epilogue -> MMA: the query tile is in TMEM
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(a_ready + 1);
    uint32_t *s_kt = tmem_slot + 1;                         // [STAGES] key tile in the stage, TCP_END = end-of-pass marker
    volatile int *s_end = reinterpret_cast<volatile int *>(s_kt + STAGES);   // [2] stream position of the pass A / pass B end marker
    volatile float *s_wthr = reinterpret_cast<volatile float *>(const_cast<int *>(s_end) + 2);   // [8] max true-distance threshold per epilogue warp
    int *s_cnt = reinterpret_cast<int *>(const_cast<float *>(s_wthr) + 8);      // [2][128] final half-list counts (-1 = overflowed)
    volatile float *s_xf = reinterpret_cast<volatile float *>(s_cnt + 2 * TC_BM); // [2][2][128] pair exchange (double-buffered)
    // [P] (bf16-truncated lower bound << 16) | tile, ascending; 8 bytes per entry while it is being sorted;
    // followed by [4][P] bf16-truncated lower bounds of the same tiles against each 32-row group of the query tile
    uint32_t *s_ord = reinterpret_cast<uint32_t *>(smem + ((reinterpret_cast<uint8_t *>(const_cast<float *>(s_xf) + 4 * TC_BM) - smem + 7) & ~(size_t)7));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long prof_start = TCP_PROF_NOW();
    // CTAs take the query tiles in the order of a.work (largest bounding box first: those scan the most key tiles,
    // and starting them last would leave the tail of the grid to a few long-running CTAs)
    const int item = a.work ? a.work[blockIdx.x] : (int)blockIdx.x;
    if (a.structured[item / a.qtiles] == 0) return;       // no low-dimensional structure: the full scan has this cloud
    const int b = item / a.qtiles;
    const int qt = item - b * a.qtiles;
    const int q0 = qt * TC_BM;
    const int tiles = a.tiles;
    const int pre = a.pre;

    if (warp == 0 && lane == 0) {
        if constexpr (!XYZ) tma_prefetch_desc(&tmap_k);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < ACC; ++s) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 8); }
        mbar_init(thr_ready, 8);
        mbar_init(a_ready, 4);
        s_end[0] = 0x7fffffff;
        s_end[1] = 0x7fffffff;
        for (int s = 0; s < 8; ++s) s_wthr[s] = CUDART_INF_F;
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, tcp_tmem_cols(C));
    if (threadIdx.x >= 64 && threadIdx.x < 80) {
        // constant query operand of the norm step: rows (1, 1, 1, 0, 0, 0, 0, 0 | 0 x 8) in bf16; one 8-row group serves
        // all 128 rows (stride between groups = 0)
        const int i = threadIdx.x - 64;
        reinterpret_cast<uint4 *>(sQe)[i] = i < 8 ? make_uint4(0x3F803F80u, 0x00003F80u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
        fence_proxy_async();
    }

    // ---- tile order: lower bound of every key tile against this CTA's query box, ascending
    {
        const float *bx = a.boxes + (size_t)b * tiles * 6;
        const int t0 = 2 * qt, t1 = min(t0 + 1, tiles - 1);
        float qlo[3], qhi[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            qlo[i] = fminf(bx[t0 * 6 + i], bx[t1 * 6 + i]);
            qhi[i] = fmaxf(bx[t0 * 6 + 3 + i], bx[t1 * 6 + 3 + i]);
        }
        const float slack = 2e-5f * sqrtf(a.nmax[b]);          // covers fp32 rounding of the projections
        // 64-bit keys: lower bound (exact bits) | centre distance (16 bits, orders the tiles whose boxes overlap) | tile
        unsigned long long *s_key = reinterpret_cast<unsigned long long *>(s_ord);
        for (int t = threadIdx.x; t < a.P; t += TCP_THREADS) {
            unsigned long long key = ~0ull;
            if (t < tiles) {
                float lb = 0.f, cd = 0.f;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const float tlo = bx[t * 6 + i], thi = bx[t * 6 + 3 + i];
                    float g = fmaxf(qlo[i] - thi, tlo - qhi[i]);
                    g = fmaxf(g - slack, 0.f);
                    lb = fmaf(g, g, lb);
                    const float dc = 0.5f * ((qlo[i] + qhi[i]) - (tlo + thi));
                    cd = fmaf(dc, dc, cd);
                }
                lb *= 0.999f;
                if (!(lb >= 0.f)) lb = 0.f;                      // NaN boxes never prune
                lb = fminf(lb, 3.0e38f);
                if (!(cd >= 0.f)) cd = 0.f;
                cd = fminf(cd, 3.0e38f);
                key = ((unsigned long long)__float_as_uint(lb) << 32) | (__float_as_uint(cd) & 0xffff0000u) | (uint32_t)t;
            }
            s_key[t] = key;
        }
        __syncthreads();
        if (a.P <= 512) {
            // few tiles: every thread ranks its own keys against all others (keys are unique), no barrier ladder
            uint32_t packed[3];
            int rank[3];
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int t = threadIdx.x + u * TCP_THREADS;
                rank[u] = -1;
                if (t < tiles) {
                    const unsigned long long kx = s_key[t];
                    int r = 0;
                    for (int j = 0; j < tiles; ++j) r += (s_key[j] < kx) ? 1 : 0;
                    rank[u] = r;
                    packed[u] = ((uint32_t)(kx >> 32) & 0xffff0000u) | ((uint32_t)kx & 0xffffu);
                }
            }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < 3; ++u)
                if (rank[u] >= 0) s_ord[rank[u]] = packed[u];
            __syncthreads();
        } else {
        for (int kk = 2; kk <= a.P; kk <<= 1)
            for (int j = kk >> 1; j > 0; j >>= 1) {
                for (int i = threadIdx.x; i < a.P; i += TCP_THREADS) {
                    const int l = i ^ j;
                    if (l > i) {
                        const unsigned long long x0 = s_key[i], x1 = s_key[l];
                        const bool up = (i & kk) == 0;
                        if ((x0 > x1) == up) { s_key[i] = x1; s_key[l] = x0; }
                    }
                }
                __syncthreads();
            }
        // compact in place to 32 bits per tile: (lower bound truncated to bf16 = rounded DOWN) | tile
        // (32-bit slot i overlays 64-bit slots <= i, all of which earlier rounds or this round's reads have consumed)
        for (int base = 0; base < a.P; base += TCP_THREADS) {
            const int i = base + threadIdx.x;
            uint32_t packed = 0;
            if (i < a.P) {
                const unsigned long long kx = s_key[i];
                packed = ((uint32_t)(kx >> 32) & 0xffff0000u) | ((uint32_t)kx & 0xffffu);
            }
            __syncthreads();
            if (i < a.P) s_ord[i] = packed;
            __syncthreads();
        }
        }
    }
    // ---- the same bound per 32-row group: an epilogue warp skips the tiles none of ITS rows can use
    uint16_t *s_lbw = reinterpret_cast<uint16_t *>(s_ord + 2 * a.P);
    {
        const float *bx = a.boxes + (size_t)b * tiles * 6;
        const float slack = 2e-5f * sqrtf(a.nmax[b]);
        for (int e = threadIdx.x; e < 4 * tiles; e += TCP_THREADS) {
            const int g = e / tiles, i = e - g * tiles;
            const int t = (int)(s_ord[i] & 0xffffu);
            float lb = 0.f;
            if (q0 + g * 32 < a.N) {
                const float *qb = a.boxes32 + ((size_t)b * tiles * 2 + (q0 >> 5) + g) * 6;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float gp = fmaxf(qb[c] - bx[t * 6 + 3 + c], bx[t * 6 + c] - qb[3 + c]);
                    gp = fmaxf(gp - slack, 0.f);
                    lb = fmaf(gp, gp, lb);
                }
                lb *= 0.999f;
                if (!(lb >= 0.f)) lb = 0.f;
                lb = fminf(lb, 3.0e38f);
            }
            s_lbw[g * a.P + i] = (uint16_t)(__float_as_uint(lb) >> 16);      // truncation rounds the bound DOWN
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const long long prof_sorted = TCP_PROF_NOW();
    long long *prof = a.prof ? a.prof + (size_t)item * 16 : nullptr;
    long long pw0 = 0, pw1 = 0;        // per-role wait cycles (measurement builds)

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            const uint8_t *ke_g = a.kext + (size_t)b * tiles * TCP_KEXT_BYTES;
            int stage = 0, seq = 0, nvis = 0;
            uint32_t phase = 0;
            // pass A feeds the threshold search (slot minima), pass B the candidate collection; both walk the tiles
            // nearest first and stop at the first tile whose lower bound exceeds every row's threshold
            for (int pass = 0; pass < 2; ++pass) {
                for (int i = 0; i < tiles; ++i) {
                    const uint32_t e = s_ord[i];
                    // pass B re-reads the `pre` nearest tiles whatever the bounds turn out to be, so they are
                    // requested while the epilogue is still finishing pass A; only then wait for its final bounds
                    if (pass == 1 && i == pre) { TCP_PROF_T0(); mbar_wait_backoff(thr_ready, 0); TCP_PROF_ADD(pw1); }
                    if (i >= pre) {
                        // thresholds only ever decrease: a stale (larger) value is safe
                        const float thr = fmaxf(fmaxf(fmaxf(s_wthr[0], s_wthr[1]), fmaxf(s_wthr[2], s_wthr[3])),
                                                fmaxf(fmaxf(s_wthr[4], s_wthr[5]), fmaxf(s_wthr[6], s_wthr[7])));
                        if (__uint_as_float(e & 0xffff0000u) > thr) break;     // every later tile has a larger bound
                    }
                    const int kt = (int)(e & 0xffffu);
                    { TCP_PROF_T0(); mbar_wait_backoff(&empty[stage], phase ^ 1); TCP_PROF_ADD(pw0); }
                    s_kt[stage] = (uint32_t)kt;
                    mbar_expect_tx(&full[stage], TILE_BYTES + TCP_KEXT_BYTES);
                    uint8_t *dst = sB + stage * TILE_BYTES;
#pragma unroll
                    for (int kb = 0; kb < NBLK; ++kb) tma_load_3d(dst + kb * BLK_BYTES, &tmap_k, &full[stage], kb * TC_KB, kt * TC_BN, b);
                    bulk_load_1d(sKe + stage * TCP_KEXT_BYTES, ke_g + (size_t)kt * TCP_KEXT_BYTES, TCP_KEXT_BYTES, &full[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    ++seq;
                    if (pass == 1) ++nvis;
                }
                // end-of-pass marker travels through the same barriers
                mbar_wait_backoff(&empty[stage], phase ^ 1);
                s_kt[stage] = TCP_END;
                mbar_arrive(&full[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
                ++seq;
                if (pass == 1 && tiles <= pre) mbar_wait_backoff(thr_ready, 0);   // (not yet waited for above)
            }
            if (a.visited) a.visited[(size_t)b * a.qtiles + qt] = nvis;
            if (prof) { prof[8] = pw0; prof[9] = pw1; prof[10] = seq; prof[11] = TCP_PROF_NOW() - prof_start; }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // The whole warp walks the loop converged and one elected lane issues (see elect_one_sync): the issuing thread's
        // own instruction stream was the bottleneck of this kernel -- 13 MMAs per tile, each wrapped in an election loop
        // plus two descriptor computations, ~1 300 cycles per tile against 450 cycles of tensor-pipe work.
        {
            mbar_wait(a_ready, 0);
            tc_fence_after();
            const uint64_t q_ext = make_kmajor_interleaved_desc(smem_u32(sQe), 128, 0);
            const uint64_t b_desc0 = make_kmajor_sw128_desc(smem_u32(sB));
            const uint64_t k_desc0 = make_kmajor_interleaved_desc(smem_u32(sKe), 128, 256);
            int stage = 0, acc = 0, seq = 0, markers = 0;
            uint32_t phase = 0, accphase = 0;
            for (;; ++seq) {
                { TCP_PROF_T0(); mbar_wait_backoff(&t_empty[acc], accphase ^ 1); TCP_PROF_ADD(pw0); }
                { TCP_PROF_T0(); mbar_wait_backoff(&full[stage], phase); TCP_PROF_ADD(pw1); }
                if (s_kt[stage] == TCP_END) {
                    if (elect_one_sync()) {
                        s_end[markers] = seq;                    // visible to the epilogue through the arrive below
                        mbar_arrive(&empty[stage]);
                        mbar_arrive(&t_full[acc]);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    if (++acc == ACC) { acc = 0; accphase ^= 1; }
                    if (++markers == 2) break;
                    continue;
                }
                tc_fence_after();
                if constexpr (XYZ) {
                    if (elect_one_sync()) {
                        // products and norm in one K = 16 step: A = the query rows in TMEM columns [0, 8)
                        umma_bf16_ts(tmem_base + ACOLS + acc * TC_BN, tmem_base, desc_advance(k_desc0, stage * TCP_KEXT_BYTES), kIdesc, 0u);
                        umma_commit(&empty[stage]);
                        umma_commit(&t_full[acc]);
                    }
                } else
                if (elect_one_sync()) {
                    const uint64_t b_desc = desc_advance(b_desc0, stage * TILE_BYTES);
                    const uint32_t d_tmem = tmem_base + ACOLS + acc * TC_BN;
#pragma unroll
                    for (int hb = 0; hb < NH; ++hb) {
#pragma unroll
                        for (int ks = 0; ks < TC_KB / 16; ++ks) {
                            // A from TMEM: a K = 16 step is 8 columns; hi in columns [0, C/2), lo in [C/2, C)
                            const uint32_t a_hi = tmem_base + hb * (TC_KB / 2) + ks * 8;
                            const uint32_t a_lo = a_hi + C / 2;
                            const uint64_t b_hi = desc_advance(b_desc, hb * BLK_BYTES + ks * 32);
                            const uint64_t b_lo = desc_advance(b_desc, (NH + hb) * BLK_BYTES + ks * 32);
                            umma_bf16_ts(d_tmem, a_hi, b_hi, kIdesc, (hb | ks) ? 1u : 0u);
                            umma_bf16_ts(d_tmem, a_hi, b_lo, kIdesc, 1);
                            umma_bf16_ts(d_tmem, a_lo, b_hi, kIdesc, 1);
                        }
                    }
                    umma_bf16(d_tmem, q_ext, desc_advance(k_desc0, stage * TCP_KEXT_BYTES), kIdesc, 1);
                    umma_commit(&empty[stage]);
                    umma_commit(&t_full[acc]);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
                if (++acc == ACC) { acc = 0; accphase ^= 1; }
            }
            if (prof && lane == 0) { prof[12] = pw0; prof[13] = pw1; prof[14] = TCP_PROF_NOW() - prof_start; }
        }
    } else {
        // ===================== epilogue: two threads per query row, 32 key columns each =====================
        // warps 2..5 take columns 0..31 of every tile, warps 6..9 columns 32..63; a warp may only touch the TMEM
        // lane quarter warp % 4, which also fixes its 32 rows
        const int ewi = warp - 2;                             // 0..7
        const int lg = warp & 3;                              // TMEM lane quarter
        const int hf = ewi >> 2;                              // column half
        const int row = lg * 32 + lane;
        const int q = q0 + row;                               // sorted position
        const bool active = q < a.N;
        const size_t srow = (size_t)b * a.N + (active ? q : 0);
        uint2 *buf = a.cand + srow * TCP_CAP;
        const float qn = active ? a.norm_pad[(size_t)b * a.Npad + q] : 0.f;
        const float margin = TC_MARGIN * sqrtf(qn * a.nmax[b]);
        float thr = active ? CUDART_INF_F : -CUDART_INF_F;
        int seq = 0, acc = 0;
        uint32_t accphase = 0;
        int xbuf = 0;                                         // exchange buffer parity (same sequence in both threads of a row)
        const uint16_t *lbw = s_lbw + lg * a.P;               // this warp's lower bound of the i-th tile of the order
        float wbound = CUDART_INF_F;                          // largest true-distance bound among this warp's rows
        const int pair_bar = 1 + lg;                          // named barrier shared by the two warps of a lane quarter

        // the query tile: one thread per row (the column-half-0 warps; a warp owns the TMEM lanes of its quarter) reads
        // the row's hi | lo operand (2C bf16), multiplies by -2 (exact: a sign and an exponent step) so that the MMAs
        // accumulate -2 q.k, and stores it into TMEM columns [0, C): one 32-bit column = two consecutive bf16 along K
        if (hf == 0 && XYZ) {
            // the row's own key image (hi | lo | hi | norm) rearranged into -2 (hi, hi, lo), 1, 1, 1, 0 ...
            const __nv_bfloat162 m2 = __float2bfloat162_rn(-2.f);
            uint32_t w[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            if (active) {
                const uint8_t *img = a.kext + ((size_t)b * tiles + (q >> 6)) * TCP_KEXT_BYTES;
                uint4 t = __ldg(reinterpret_cast<const uint4 *>(img + tcp_kext_offset(q & 63, 0)));
                __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&t);
#pragma unroll
                for (int e = 0; e < 3; ++e) h[e] = __hmul2(h[e], m2);
                const uint32_t n0 = t.x, n1 = t.y, n2 = t.z;      // -2 (h0, h1), -2 (h2, l0), -2 (l1, l2)
                w[0] = n0;
                w[1] = (n1 & 0xffffu) | (n0 << 16);                // -2 (h2, h0)
                w[2] = (n0 >> 16) | (n1 << 16);                    // -2 (h1, h2)
                w[3] = (n1 >> 16) | (n2 << 16);                    // -2 (l0, l1)
                w[4] = (n2 >> 16) | 0x3F800000u;                   // (-2 l2, 1)
                w[5] = 0x3F803F80u;                                // (1, 1)
            }
            tmem_st8(tmem_base + ((uint32_t)(lg * 32) << 16), w);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready);
        }
        if (hf == 0 && !XYZ) {
            const __nv_bfloat162 m2 = __float2bfloat162_rn(-2.f);
            const uint4 *src = reinterpret_cast<const uint4 *>(a.xs + ((size_t)b * a.N + (active ? q : 0)) * 2 * C);
#pragma unroll
            for (int blk = 0; blk < ACOLS / 32; ++blk) {
                uint32_t w[32];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    uint4 t = active ? __ldg(src + blk * 8 + u) : make_uint4(0u, 0u, 0u, 0u);
                    __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&t);
#pragma unroll
                    for (int e = 0; e < 4; ++e) h[e] = __hmul2(h[e], m2);
                    w[u * 4 + 0] = t.x; w[u * 4 + 1] = t.y; w[u * 4 + 2] = t.z; w[u * 4 + 3] = t.w;
                }
                tmem_st32(tmem_base + ((uint32_t)(lg * 32) << 16) + blk * 32, w);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready);
        }
        const long long prof_scaled = TCP_PROF_NOW();
        long long prof_refresh = 0;

        // the two threads of a row combine a value (sum for counts, min / max for ranges) through shared memory
        auto pair_exchange = [&](float mine) -> float {
            s_xf[(xbuf * 2 + hf) * TC_BM + row] = mine;
            named_bar_sync(pair_bar, 64);
            const float other = s_xf[(xbuf * 2 + (hf ^ 1)) * TC_BM + row];
            xbuf ^= 1;
            return other;
        };
        // smallest bound with count(slot minima of the row <= bound) >= k, by bisection over both threads' 32 slots
        auto row_bound = [&](const float (&m)[32 * SM], int iters) -> float {
            float mn = CUDART_INF_F, mx = -CUDART_INF_F;
            int nf = 0;
#pragma unroll
            for (int s = 0; s < 32 * SM; ++s) {
                mn = fminf(mn, m[s]);
                if (m[s] < CUDART_INF_F) { mx = fmaxf(mx, m[s]); ++nf; }
            }
            mn = fminf(mn, pair_exchange(mn));
            mx = fmaxf(mx, pair_exchange(mx));
            nf += (int)pair_exchange((float)nf);
            float lo = mn, hi = mx;
            int c_hi = nf;
            // The two warps of a lane quarter hold the same 32 rows and see the same totals, so they leave the loop in
            // the same iteration (the exchanges stay aligned): when every row has exactly k slots below its bound
            // (nothing left to gain) or after `iters` halvings.
            for (int it = 0; it < iters; ++it) {
                if (__all_sync(FULLW, c_hi <= a.k)) break;
                const float mid = 0.5f * lo + 0.5f * hi;
                int c = 0;
#pragma unroll
                for (int s = 0; s < 32 * SM; ++s) c += (m[s] <= mid) ? 1 : 0;
                c += (int)pair_exchange((float)c);
                if (c >= a.k) { hi = mid; c_hi = c; } else lo = mid;
            }
            return nf >= a.k ? hi : CUDART_INF_F;
        };

        // ---- pass A: per-column-slot minima over the visited tiles.  The 64 slot minima of a row belong to 64
        // different keys (SM = 2: the two smallest per slot, 128 keys), so the k-th smallest of them bounds the
        // row's k-th distance from above; nothing is stored.
        {
            float m[32 * SM];
#pragma unroll
            for (int s = 0; s < 32 * SM; ++s) m[s] = CUDART_INF_F;
            int done = 0, refresh_at = a.refresh0;
            for (;; ++seq) {
                { TCP_PROF_T0(); mbar_wait(&t_full[acc], accphase); TCP_PROF_ADD(pw0); }
                if (seq == s_end[0]) break;
                tc_fence_after();
                // a tile whose bound against this warp's 32 rows exceeds all their current bounds holds no key that
                // could lower any of them: its slot minima are not needed (the final bound is the same without them)
                if (__uint_as_float((uint32_t)lbw[seq] << 16) <= wbound) {
                    const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + ACOLS + acc * TC_BN + hf * 32;
                    if constexpr (SM == 1) {
                        uint32_t v[32];
                        tmem_ld32(taddr, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int s = 0; s < 32; ++s) m[s] = fminf(m[s], __uint_as_float(v[s]));     // |k|^2 - 2 q.k, +inf past N
                    } else {
                        // 64 slot registers: the accumulator chunk is read 16 columns at a time to stay within the
                        // register budget of two CTAs per SM
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            uint32_t v[16];
                            tmem_ld16(taddr + h * 16, v);
                            tmem_ld_wait();
#pragma unroll
                            for (int e = 0; e < 16; ++e) {
                                const int sl = h * 16 + e;
                                const float d = __uint_as_float(v[e]);
                                m[32 + sl] = fminf(m[32 + sl], fmaxf(m[sl], d));      // second smallest of the slot
                                m[sl] = fminf(m[sl], d);
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&t_empty[acc]);
                if (++acc == ACC) { acc = 0; accphase ^= 1; }
                if (++done == refresh_at) {
                    // let the producer start skipping: publish the bound reached so far (after 12 and 36 tiles: each
                    // refresh costs about as much as six tiles of this pass)
                    refresh_at = refresh_at < a.refresh_max ? refresh_at * a.refresh_mul : 0x7fffffff;
                    TCP_PROF_T0();
                    const float bd = row_bound(m, 6);
                    TCP_PROF_ADD(prof_refresh);
                    float wt = active ? bd + margin + qn : -CUDART_INF_F;
                    for (int o = 16; o; o >>= 1) wt = fmaxf(wt, __shfl_xor_sync(FULLW, wt, o));
                    if (lane == 0) s_wthr[ewi] = wt;
                    wbound = wt;
                }
            }
            // end marker of pass A: final thresholds (bit-identical in the two threads of a row)
            const long long prof_a_end = TCP_PROF_NOW();
            const float bd = row_bound(m, 11);
            if (active) thr = bd + margin;
            float wt = active ? thr + qn : -CUDART_INF_F;
            for (int o = 16; o; o >>= 1) wt = fmaxf(wt, __shfl_xor_sync(FULLW, wt, o));
            __syncwarp();
            if (lane == 0) {
                s_wthr[ewi] = wt;
                mbar_arrive(thr_ready);
                mbar_arrive(&t_empty[acc]);
            }
            if (++acc == ACC) { acc = 0; accphase ^= 1; }
            ++seq;
            wbound = wt;
            if (prof && ewi == 0 && lane == 0) {
                prof[0] = prof_sorted - prof_start; prof[1] = prof_scaled - prof_sorted; prof[2] = prof_a_end - prof_scaled;
                prof[3] = TCP_PROF_NOW() - prof_a_end; prof[4] = pw0; prof[5] = prof_refresh;
            }
            pw0 = 0;
        }
        const long long prof_b_start = TCP_PROF_NOW();

        // ---- pass B: same order again, every key below the row's threshold becomes a candidate.  Each thread owns
        // half of the row's list (capacity TCP_HCAP): no shared counter, and a half that fills up is compacted by
        // its warp exactly like in the full scan (its k-th smallest entry is a valid bound for the whole row).
        uint2 *hbuf = buf + hf * TCP_HCAP;
        // The half-list is 2 KB and 2 KB-aligned, so appending only ever changes the low 32 address bits: the write
        // pointer is kept as a (running low word, constant high word) pair -- one predicated add per append instead
        // of a 64-bit index computation.
        const uint32_t hb_lo0 = (uint32_t)reinterpret_cast<uintptr_t>(hbuf);
        unsigned long long wp = reinterpret_cast<unsigned long long>(hbuf);
        int cnt = 0;
        bool ovf = false;
        for (int i = 0;; ++i, ++seq) {
            { TCP_PROF_T0(); mbar_wait(&t_full[acc], accphase); TCP_PROF_ADD(pw0); }
            if (seq == s_end[1]) break;
            tc_fence_after();
            if (__uint_as_float((uint32_t)lbw[i] << 16) > wbound) {
                // no row of this warp can have a candidate in the tile: only keep the pipeline moving
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&t_empty[acc]);
                if (++acc == ACC) { acc = 0; accphase ^= 1; }
                continue;
            }
            const int kt = (int)(s_ord[i] & 0xffffu);
            const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + ACOLS + acc * TC_BN + hf * 32;
            uint32_t v[32];
            tmem_ld32(taddr, v);
            tmem_ld_wait();
            const int jbase = kt * TC_BN + hf * 32;
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float d = __uint_as_float(v[c4 * 4 + e]);                      // |k|^2 - 2 q.k, +inf for keys >= N
                    // if (d < thr) { *wp = (d, index); ++wp; } as two predicated instructions: only the low address
                    // word ever changes (see above)
                    asm volatile(
                        "{\n"
                        ".reg .pred p;\n"
                        ".reg .u32 lo, hi;\n"
                        "setp.lt.f32 p, %1, %2;\n"
                        "@p st.global.v2.b32 [%0], {%3, %4};\n"
                        "mov.b64 {lo, hi}, %0;\n"
                        "@p add.u32 lo, lo, 8;\n"
                        "mov.b64 %0, {lo, hi};\n"
                        "}\n"
                        : "+l"(wp)
                        : "f"(d), "f"(thr), "r"(__float_as_uint(d)), "r"(jbase + c4 * 4 + e)
                        : "memory");
                }
            }
            cnt = (int)(((uint32_t)wp - hb_lo0) >> 3);
            // accumulator drained: hand it back before any list maintenance
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[acc]);
            if (++acc == ACC) { acc = 0; accphase ^= 1; }
            // half-lists that could overflow during the next 32 columns are compacted now (rare)
            unsigned need = __ballot_sync(FULLW, cnt > TCP_HCAP - 32);
            while (need) {
                const int r = __ffs(need) - 1;
                need &= need - 1;
                const int n_r = __shfl_sync(FULLW, cnt, r);
                const float m_r = __shfl_sync(FULLW, margin, r);
                const unsigned long long p_r = __shfl_sync(FULLW, (unsigned long long)hbuf, r);
                __syncwarp();
                int nc; float nt;
                compact_row<TCP_HCAP>(reinterpret_cast<uint2 *>(p_r), n_r, a.k, m_r, lane, nc, nt);
                if (lane == r) {
                    cnt = nc;
                    thr = fminf(thr, nt);
                    if (nc > TCP_HCAP - 32) { ovf = true; cnt = 0; thr = -CUDART_INF_F; }
                    wp = reinterpret_cast<unsigned long long>(hbuf + cnt);
                }
            }
        }

        if (prof && ewi == 0 && lane == 0) { prof[6] = TCP_PROF_NOW() - prof_b_start; prof[7] = pw0; }
        // the row's two halves meet through shared memory: both must be done before the totals are written
        s_cnt[hf * TC_BM + row] = ovf ? -1 : cnt;
        named_bar_sync(5, 256);
        if (active && hf == 0) {
            const int c0 = s_cnt[row], c1 = s_cnt[TC_BM + row];
            const bool bad = c0 < 0 || c1 < 0 || c0 + c1 < a.k;
            a.cand_cnt[2 * srow] = bad ? 0 : c0;
            a.cand_cnt[2 * srow + 1] = bad ? 0 : c1;
            a.overflow[(size_t)b * a.N + a.perm[srow]] = bad ? 1 : 0;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tcp_tmem_cols(C));
    }
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

static size_t tcp_cub_temp_bytes(size_t n, int end_bit) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned *)nullptr, (unsigned *)nullptr,
                                    (const int *)nullptr, (int *)nullptr, (int)n, 0, end_bit);
    return bytes;
}
static int tcp_cloud_bits(int B) {
    int bits = 0;
    while ((1 << bits) < B) ++bits;
    return bits;
}
// Morton bits per axis so that (cloud | code) fits a 32-bit radix key
static int tcp_axis_bits(int B) {
    // 24 key bits = three 8-bit radix passes; 6 bits per axis (262 144 cells) at B = 16 is far finer than a 64-point tile
    int bits = (24 - tcp_cloud_bits(B)) / 3;
    if (bits < 5) bits = (32 - tcp_cloud_bits(B)) / 3;      // many clouds: spend a fourth pass rather than coarsen the cells
    return bits > 10 ? 10 : bits;
}
static bool tcp_supported(int B, int N, int k2) {
    return N >= TCP_MIN_N && k2 <= 2 * TC_BN && ceil_div(N, TC_BN) <= 2048 && tcp_axis_bits(B) >= 5;
}

size_t knn_tc_workspace_bytes(int B, int C, int N) {
    size_t bn = (size_t)B * N;
    size_t t = 0;
    // pruned path
    t += align_up((size_t)B * TCP_SPLIT * (C * C + C) * sizeof(float));   // partial moments
    t += align_up((size_t)B * (3 * C + 4) * sizeof(float));               // pca (xyz clouds: the bounding boxes, 8 floats each)
    t += align_up((size_t)B * sizeof(int));                               // structured flags
    t += align_up(3 * bn * sizeof(float));                                // projections
    t += 2 * align_up(bn * sizeof(unsigned));                             // keys in/out
    t += 2 * align_up(bn * sizeof(int));                                  // vals in/out
    t += align_up(tcp_cub_temp_bytes(bn, 32));                            // cub temp
    t += 2 * align_up(bn * sizeof(int));                                  // perm, inv
    t += 3 * align_up((size_t)B * ceil_div(N, TC_BN) * 6 * sizeof(float));   // tile boxes + boxes of their 32-point halves
    t += align_up((size_t)B * ceil_div(N, TC_BM) * sizeof(int));          // visited-tile statistics
    t += 2 * align_up((size_t)B * ceil_div(N, TC_BM) * sizeof(int));      // launch order of the query tiles + its sort keys
    t += align_up((size_t)B * ceil_div(N, TC_BM) * 16 * sizeof(long long)); // per-CTA cycle counters (measurement builds)
    t += 2 * align_up(bn * sizeof(int)) + align_up(2 * (size_t)B * sizeof(int));  // fallback / wide re-rank row lists + counts
    t += align_up(bn * 2 * C * sizeof(__nv_bfloat16));   // xs
    t += align_up(bn * (C == 3 ? 4 : C) * sizeof(float)); // x_nc (xyz clouds: rows padded to 16 bytes)
    t += align_up(bn * sizeof(float));                   // norm
    t += align_up((size_t)B * (ceil_div(N, TC_BN) * TC_BN) * sizeof(float));   // norm_pad
    t += align_up((size_t)B * ceil_div(N, TC_BN) * TCP_KEXT_BYTES);            // key norms as an operand step
    t += align_up((size_t)B * sizeof(float));            // nmax
    t += align_up((bn * TCP_CAP + 512) * sizeof(uint2)); // cand, 4 KB-aligned rows (the full scan uses TC_CAP entries per row of it)
    t += align_up(2 * bn * sizeof(int));                 // cand_cnt (two halves per row on the pruned path)
    t += align_up(bn * sizeof(int));                     // overflow
    return t;
}

bool knn_tc_supported(int C, int N, int k2) {
    return (C == 64 || C == 128) && k2 <= 128 && N >= TC_BM && k2 + TC_SLACK + 64 <= TC_CAP;
}
// xyz clouds (C = 3, L2): only the box-pruned scan exists for them (one K = 16 MMA per key tile); everything it does not
// cover stays with knn_xyz.cu
// (and the largest clouds: at 4 x 100 000 points the two are level, 2.07 vs 2.00 ms, and its workspace is 40x smaller)
bool knn_tc_xyz_supported(int B, int N, int k2) {
    return tcp_supported(B, N, k2) && N < TCP_FULL_MIN_N && k2 + TC_SLACK + 64 <= TC_CAP;
}

// knn_xyz.cu: bbox[b][8], keys = (cloud << 3 ab) | Morton code with `ab` bits per axis, vals = point index
int launch_xyz_sort_keys(const float *x, float *bbox, unsigned *keys, int *vals, int B, int C, int N, int ab, cudaStream_t st);

// declared in knn_select.cu
int launch_sqnorm_public(const float *x, float *out, int B, int C, int Cuse, int N, cudaStream_t st);
int knn_fallback_rows(const float *x, const float *norms, const int *qlist, const int *qcount, int B, int C, int N, int k1, int k2,
                      int64_t *idx64, int32_t *idx32, cudaStream_t st);

template <int C, int MODE = 0>
static int launch_tc(const CUtensorMap &tmap_q, const CUtensorMap &tmap_k, TcScanArgs sa, RerankArgs ra, int B, cudaStream_t st) {
    constexpr int NBLK = 2 * C / TC_KB;
    constexpr int STAGES = 2;
    const size_t smem = 1024 + (size_t)NBLK * TC_BM * 128 + (size_t)STAGES * NBLK * TC_BN * 128 +
                        TC_NRING * TC_BN * sizeof(float) + 32 * sizeof(uint64_t);
    auto kern = knn_tc_scan_kernel<C, MODE>;
    GCANET_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(sa.N, TC_BM), B);
    probe_scan_begin(st);
    kern<<<grid, TC_THREADS, smem, st>>>(tmap_q, tmap_k, sa);
    GCANET_LAUNCH_OK("knn_tc_scan_kernel");
    probe_scan_end(st);
    if (sa.debug_no_append) return GCANET_OK;      // measurement aid: scan pipeline only
    dim3 rgrid(ceil_div(sa.N, 8), B);
    knn_tc_rerank_kernel<C, false><<<rgrid, 256, 0, st>>>(ra);
    GCANET_LAUNCH_OK("knn_tc_rerank_kernel");
    return GCANET_OK;
}

template <int C, int SM>
static int launch_tcp(const CUtensorMap &tmap_q, const CUtensorMap &tmap_k, TcpScanArgs sa, TcScanArgs fa, RerankArgs ra, int B,
                      cudaStream_t st) {
    constexpr int NBLK = C == 3 ? 0 : 2 * C / TC_KB;
    constexpr int STAGES = tcp_stages(C);
    const size_t smem = 1024 + (size_t)STAGES * NBLK * TC_BN * 128 +
                        (size_t)STAGES * TCP_KEXT_BYTES + 256 + 32 * sizeof(uint64_t) + 6 * TC_BM * sizeof(float) + (size_t)sa.P * 16;
    auto kern = knn_tcp_scan_kernel<C, SM>;
    GCANET_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(sa.N, TC_BM) * B);
    probe_scan_begin(st);
    kern<<<grid, TCP_THREADS, smem, st>>>(tmap_k, sa);
    GCANET_LAUNCH_OK("knn_tcp_scan_kernel");
    if (C == 3) probe_scan_end(st);
    if constexpr (C != 3) {
        // clouds without low-dimensional structure (their sorted order is the original order): single-pass full scan
        constexpr int FS_STAGES = 2;
        const size_t fsmem = 1024 + (size_t)NBLK * TC_BM * 128 + (size_t)FS_STAGES * NBLK * TC_BN * 128 +
                             TC_NRING * TC_BN * sizeof(float) + 32 * sizeof(uint64_t);
        auto fkern = knn_tc_scan_kernel<C, 0>;
        GCANET_CUDA_OK(cudaFuncSetAttribute(fkern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
        fkern<<<dim3(ceil_div(sa.N, TC_BM), B), TC_THREADS, fsmem, st>>>(tmap_q, tmap_k, fa);
        GCANET_LAUNCH_OK("knn_tc_scan_kernel");
        probe_scan_end(st);
    }
    dim3 rgrid(ceil_div(sa.N, 8), B);
    knn_tc_rerank_kernel<C, false><<<rgrid, 256, 0, st>>>(ra);
    GCANET_LAUNCH_OK("knn_tc_rerank_kernel");
    knn_tc_rerank_kernel<C, true><<<dim3(32, B), 256, 0, st>>>(ra);       // rows with more than TC_CAP candidates (usually none)
    GCANET_LAUNCH_OK("knn_tc_rerank_kernel<big>");
    return GCANET_OK;
}

template <int C>
static int launch_tcp_prep(const float *x, float *part, float *pca, int *structured, float *proj, unsigned *keys, int *vals,
                           float *norm, int B, int N, cudaStream_t st) {
    int stride = N / 1024;
    stride = stride < 1 ? 1 : (stride > 8 ? 8 : stride);
    const int Ns = (N + stride - 1) / stride;
    tcp_moments_kernel<C><<<dim3(TCP_SPLIT, B), 256, 0, st>>>(x, part, N, Ns, stride);
    GCANET_LAUNCH_OK("tcp_moments_kernel");
    const size_t smem = ((size_t)C * (C + 1) + 7 * C) * sizeof(float);
    auto kern = tcp_pca_kernel<C>;
    if (smem > 48 * 1024) GCANET_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<B, 256, smem, st>>>(part, pca, structured, Ns, N);
    GCANET_LAUNCH_OK("tcp_pca_kernel");
    tcp_project_kernel<C><<<dim3(ceil_div(N, 128), B), 128, 0, st>>>(x, pca, proj, keys, vals, norm, B, N, tcp_axis_bits(B));
    GCANET_LAUNCH_OK("tcp_project_kernel");
    return GCANET_OK;
}

int knn_graph_tensor_cores(const float *x, int B, int C, int N, int k1, int k2, int64_t *idx64, int32_t *idx32,
                           void *ws, int unordered, int no_prune, cudaStream_t st) {
    EncodeTiledFn encode = get_encode_fn();
    if (!encode) { set_error("knn_graph: cuTensorMapEncodeTiled is not available from the driver"); return GCANET_ERR_CUDA; }
    const size_t bn = (size_t)B * N;
    const bool xyz = C == 3;                               // xyz clouds: pruned scan only (the caller checked knn_tc_xyz_supported)
    Carver cv(ws);
    float *part = cv.take<float>((size_t)B * TCP_SPLIT * (C * C + C));
    float *pca = cv.take<float>((size_t)B * (3 * C + 4));
    int *structured = cv.take<int>(B);
    float *proj = cv.take<float>(3 * bn);
    unsigned *keys_in = cv.take<unsigned>(bn);
    unsigned *keys_out = cv.take<unsigned>(bn);
    int *vals_in = cv.take<int>(bn);
    int *vals_out = cv.take<int>(bn);
    const int end_bit = tcp_cloud_bits(B) + 3 * tcp_axis_bits(B);
    size_t temp_bytes = tcp_cub_temp_bytes(bn, 32);
    void *temp = cv.take<char>(temp_bytes);
    int *perm = cv.take<int>(bn);
    int *inv = cv.take<int>(bn);
    float *boxes = cv.take<float>((size_t)B * ceil_div(N, TC_BN) * 6);
    float *boxes32 = cv.take<float>((size_t)B * ceil_div(N, TC_BN) * 12);
    int *visited = cv.take<int>((size_t)B * ceil_div(N, TC_BM));
    int *work_buf = cv.take<int>((size_t)B * ceil_div(N, TC_BM));
    long long *prof_buf = cv.take<long long>((size_t)B * ceil_div(N, TC_BM) * 16);
    unsigned *wkey = cv.take<unsigned>((size_t)B * ceil_div(N, TC_BM));
    int *fb_list = cv.take<int>(bn);
    int *big_list = cv.take<int>(bn);
    int *fb_count = cv.take<int>(2 * (size_t)B);          // fallback rows | wide re-rank rows
    int *big_count = fb_count + B;
    const char *env_np = GCANET_AID_ENV("GCANET_TC_NO_PRUNE");
    const bool prune = xyz || (!no_prune && tcp_supported(B, N, k2) && !(env_np && env_np[0] == '1') && !GCANET_AID_ENV("GCANET_TC_DEBUG"));
    // small clouds: one CTA per cloud sorts and does the whole per-cloud preparation (tcp_cloud_prep_kernel)
    const bool cloud_prep = prune && N <= TCP_SORT_MAX_N && !GCANET_AID_ENV("GCANET_TC_DEVICE_SORT");
    const int qtiles_all = ceil_div(N, TC_BM);
    const bool want_order = prune && B * qtiles_all <= TCP_MAX_WORK && !GCANET_AID_ENV("GCANET_TC_NO_ORDER");
    if (!cloud_prep) GCANET_CUDA_OK(cudaMemsetAsync(fb_count, 0, 2 * B * sizeof(int), st));
    __nv_bfloat16 *xs = cv.take<__nv_bfloat16>(bn * 2 * C);
    float *x_nc = cv.take<float>(bn * (xyz ? 4 : C));
    float *norm = cv.take<float>(bn);
    const int Npad = ceil_div(N, TC_BN) * TC_BN;
    float *norm_pad = cv.take<float>((size_t)B * Npad);
    uint8_t *kext = cv.take<uint8_t>((size_t)B * ceil_div(N, TC_BN) * TCP_KEXT_BYTES);
    float *nmax = cv.take<float>(B);
    // rows of the candidate buffer start on 4 KB boundaries whatever the caller's workspace alignment: the scan
    // advances a half-list's write pointer in the low address word only, so a half-list (2 KB, 2 KB-aligned)
    // must never straddle a 4 GB boundary
    uint2 *cand = cv.take<uint2>(bn * TCP_CAP + 512);
    cand = reinterpret_cast<uint2 *>((reinterpret_cast<uintptr_t>(cand) + 4095) & ~(uintptr_t)4095);
    int *cand_cnt = cv.take<int>(2 * bn);
    int *overflow = cv.take<int>(bn);

    int rc = GCANET_OK;
    const int tiles = ceil_div(N, TC_BN);
    TcpTileArgs ta{vals_out, norm, xyz ? x : proj, perm, inv, norm_pad, boxes, boxes32, reinterpret_cast<unsigned *>(nmax), kext,
                   B, N, Npad, tiles, C, reinterpret_cast<float4 *>(x_nc), structured};
    TcpCloudPrepArgs ca{ta, keys_in, norm, vals_out, fb_count, big_count, 3 * tcp_axis_bits(B), tcp_axis_bits(B)};
    // squared norms: the projection kernel (feature clouds) and the per-cloud kernel (xyz clouds) compute them on the way
    if (!prune || (xyz && !cloud_prep)) {
        rc = launch_sqnorm_public(x, norm, B, C, C, N, st);
        if (rc) return rc;
    }
    if (prune && !xyz) {
        rc = C == 64 ? launch_tcp_prep<64>(x, part, pca, structured, proj, keys_in, vals_in, norm, B, N, st)
                     : launch_tcp_prep<128>(x, part, pca, structured, proj, keys_in, vals_in, norm, B, N, st);
        if (rc) return rc;
    }
    if (cloud_prep) {
        if (xyz) tcp_cloud_prep_kernel<true><<<B, 1024, 0, st>>>(ca);
        else tcp_cloud_prep_kernel<false><<<B, 1024, 0, st>>>(ca);
        GCANET_LAUNCH_OK("tcp_cloud_prep_kernel");
        if (xyz) tcp_tiles_kernel<true><<<dim3(ceil_div(tiles, 8), B), 256, 0, st>>>(ta);
        else tcp_tiles_kernel<false><<<dim3(ceil_div(tiles, 8), B), 256, 0, st>>>(ta);
        GCANET_LAUNCH_OK("tcp_tiles_kernel");
    } else if (prune) {
        if (xyz) {
            // sort key = (cloud, curve position of the coordinates inside the cloud's bounding box); `pca` holds the boxes
            rc = launch_xyz_sort_keys(x, pca, keys_in, vals_in, B, C, N, tcp_axis_bits(B), st);
            if (rc) return rc;
        }
        GCANET_CUDA_OK(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, vals_in, vals_out, (int)bn, 0, end_bit, st));
        count_launch();
        GCANET_CUDA_OK(cudaMemsetAsync(nmax, 0, B * sizeof(float), st));
        if (xyz) tcp_tiles_kernel<true><<<dim3(ceil_div(tiles, 8), B), 256, 0, st>>>(ta);
        else tcp_tiles_kernel<false><<<dim3(ceil_div(tiles, 8), B), 256, 0, st>>>(ta);
        GCANET_LAUNCH_OK("tcp_tiles_kernel");
    }
    if (!xyz) {
        const uintptr_t bases = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(xs) | reinterpret_cast<uintptr_t>(x_nc);
        if (N % 4 == 0 && C % 64 == 0 && (bases & 15) == 0) {
            tc_prep_wide_kernel<<<dim3(ceil_div(N, 64), C / 64, B), 256, 0, st>>>(x, xs, x_nc, prune ? inv : nullptr, C, N);
            GCANET_LAUNCH_OK("tc_prep_wide_kernel");
        } else {
            dim3 grid(ceil_div(N, 32), ceil_div(C, 32), B), block(32, 8);
            tc_prep_kernel<<<grid, block, 0, st>>>(x, xs, x_nc, prune ? inv : nullptr, C, N);
            GCANET_LAUNCH_OK("tc_prep_kernel");
        }
        if (!prune) {
            tc_normmax_kernel<<<B, 256, 0, st>>>(norm, nmax, norm_pad, N, Npad);
            GCANET_LAUNCH_OK("tc_normmax_kernel");
        }
    }

    // 3-D tensor map over xs: (K = 2C bf16, N rows, B clouds), box = (64, 128, 1), 128-byte swizzle;
    // rows past N are zero-filled, so a partial last tile never reads the next cloud.
    CUtensorMap tmap_q, tmap_k;
    memset(&tmap_q, 0, sizeof(tmap_q));
    memset(&tmap_k, 0, sizeof(tmap_k));
    cuuint64_t gdim[3] = {(cuuint64_t)(2 * C), (cuuint64_t)N, (cuuint64_t)B};
    cuuint64_t gstride[2] = {(cuuint64_t)(2 * C) * sizeof(__nv_bfloat16), (cuuint64_t)N * 2 * C * sizeof(__nv_bfloat16)};
    cuuint32_t estride[3] = {1, 1, 1};
    cuuint32_t box_q[3] = {(cuuint32_t)TC_KB, (cuuint32_t)TC_BM, 1};
    cuuint32_t box_k[3] = {(cuuint32_t)TC_KB, (cuuint32_t)TC_BN, 1};
    CUresult cr = CUDA_SUCCESS;                            // (xyz clouds need no tensor map: their operand is one bulk copy per tile)
    if (!xyz)
        cr = encode(&tmap_q, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, xs, gdim, gstride, box_q, estride,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (!xyz && cr == CUDA_SUCCESS)
        cr = encode(&tmap_k, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, xs, gdim, gstride, box_k, estride,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { set_error("knn_graph: cuTensorMapEncodeTiled failed (%d)", (int)cr); return GCANET_ERR_CUDA; }

    if (prune) {
        int P = 16;
        while (P < tiles) P <<= 1;
        int pre = TCP_PRE;
        if (const char *e = GCANET_AID_ENV("GCANET_TC_PRE")) pre = atoi(e);          // measurement aid
        if (pre > tiles) pre = tiles;
        if (pre < 0) pre = 0;
        const int qtiles = ceil_div(N, TC_BM);
        if (GCANET_AID_ENV("GCANET_TC_STATS")) GCANET_CUDA_OK(cudaMemsetAsync(visited, 0, (size_t)B * qtiles * sizeof(int), st));
        const int *work = nullptr;
        if (want_order) {
            tcp_work_key_kernel<<<dim3(ceil_div(qtiles, 8), B), 256, 0, st>>>(boxes, wkey, tiles, qtiles);
            GCANET_LAUNCH_OK("tcp_work_key_kernel");
            tcp_work_order_kernel<<<ceil_div(B * qtiles, 8), 256, 0, st>>>(wkey, work_buf, B * qtiles);
            GCANET_LAUNCH_OK("tcp_work_order_kernel");
            work = work_buf;
        }
        long long *prof = GCANET_AID_ENV("GCANET_TC_PROF") ? prof_buf : nullptr;
        if (prof) GCANET_CUDA_OK(cudaMemsetAsync(prof, 0, (size_t)B * qtiles * 16 * sizeof(long long), st));
        int r0 = 12, rmul = 3, rmax = 36;
        if (const char *e = GCANET_AID_ENV("GCANET_TC_REFRESH")) sscanf(e, "%d,%d,%d", &r0, &rmul, &rmax);   // measurement aid
        TcpScanArgs sa{norm_pad, Npad, xs, kext, nmax, boxes, boxes32, perm, cand, cand_cnt, overflow, visited, work, structured, N, k2, tiles, pre, P, qtiles, prof,
                       r0, rmul, rmax};
        RerankArgs ra{x_nc, norm, nmax, cand, cand_cnt, overflow, idx64, idx32, N, k2, k2 / k1, gcanet_knn_graph_columns(k1, k2),
                      (unordered && k1 == k2) ? 1 : 0, perm, TCP_CAP, 1, big_list, big_count, fb_list, fb_count};
        int fstride = (int)(tiles * 0.381966f);
        if (fstride < 1) fstride = 1;
        auto gcd2 = [](int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; };
        while (gcd2(fstride, tiles) != 1) ++fstride;
        TcScanArgs fa{norm_pad, Npad, norm, nmax, cand, cand_cnt, overflow, N, k2, tiles, 0, fstride, structured, TCP_CAP, 1};
        if (xyz) rc = k2 <= TC_BN ? launch_tcp<3, 1>(tmap_q, tmap_k, sa, fa, ra, B, st) : launch_tcp<3, 2>(tmap_q, tmap_k, sa, fa, ra, B, st);
        else if (k2 <= TC_BN && !GCANET_AID_ENV("GCANET_TC_SM2")) rc = C == 64 ? launch_tcp<64, 1>(tmap_q, tmap_k, sa, fa, ra, B, st) : launch_tcp<128, 1>(tmap_q, tmap_k, sa, fa, ra, B, st);
        else rc = C == 64 ? launch_tcp<64, 2>(tmap_q, tmap_k, sa, fa, ra, B, st) : launch_tcp<128, 2>(tmap_q, tmap_k, sa, fa, ra, B, st);
        if (rc) return rc;
        if (prof) {                                // measurement aid: synchronises; mean cycles per CTA and role
            const int nq = B * qtiles;
            long long *h = (long long *)malloc((size_t)nq * 16 * sizeof(long long));
            GCANET_CUDA_OK(cudaStreamSynchronize(st));
            GCANET_CUDA_OK(cudaMemcpy(h, prof, (size_t)nq * 16 * sizeof(long long), cudaMemcpyDeviceToHost));
            double m[16] = {0};
            for (int i = 0; i < nq; ++i) for (int j = 0; j < 16; ++j) m[j] += (double)h[(size_t)i * 16 + j] / nq;
            fprintf(stderr, "[gcanet] scan CTA cycles (mean of %d): prologue %.0f | A-scale %.0f | pass A %.0f (wait t_full %.0f, refresh %.0f) | "
                    "final bound %.0f | pass B %.0f (wait t_full %.0f) || producer: total %.0f, wait empty %.0f, wait thr %.0f, steps %.1f || "
                    "mma: total %.0f, wait t_empty %.0f, wait full %.0f\n",
                    nq, m[0], m[1], m[2], m[4], m[5], m[3], m[6], m[7], m[11], m[8], m[9], m[10], m[14], m[12], m[13]);
            free(h);
        }
        const char *stats = GCANET_AID_ENV("GCANET_TC_STATS");
        if (stats && stats[0] == '1') {            // measurement aid: synchronises (clouds left to the full scan report 0 tiles)
            const int nq = B * ceil_div(N, TC_BM);
            int *h = (int *)malloc(nq * sizeof(int));
            GCANET_CUDA_OK(cudaStreamSynchronize(st));
            GCANET_CUDA_OK(cudaMemcpy(h, visited, nq * sizeof(int), cudaMemcpyDeviceToHost));
            double s = 0;
            for (int i = 0; i < nq; ++i) s += h[i];
            qsort(h, nq, sizeof(int), [](const void *u, const void *v) { return *(const int *)u - *(const int *)v; });
            fprintf(stderr, "[gcanet] pruned kNN C=%d: %.1f of %d key tiles visited per query tile (min %d, median %d, p90 %d, p99 %d, max %d)\n",
                    C, s / nq, tiles, h[0], h[nq / 2], h[nq * 9 / 10], h[nq * 99 / 100], h[nq - 1]);
            free(h);
            int *hc2 = (int *)malloc(2 * bn * sizeof(int)), *hc = (int *)malloc(bn * sizeof(int)), *ho = (int *)malloc(bn * sizeof(int));
            GCANET_CUDA_OK(cudaMemcpy(hc2, cand_cnt, 2 * bn * sizeof(int), cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < bn; ++i) hc[i] = hc2[2 * i] + hc2[2 * i + 1];
            free(hc2);
            GCANET_CUDA_OK(cudaMemcpy(ho, overflow, bn * sizeof(int), cudaMemcpyDeviceToHost));
            double sc = 0;
            long no = 0, nz = 0, big = 0;
            int mxc = 0;
            for (size_t i = 0; i < bn; ++i) { sc += hc[i]; no += ho[i]; nz += hc[i] == 0; big += hc[i] > 160; if (hc[i] > mxc) mxc = hc[i]; }
            fprintf(stderr, "[gcanet]   candidates per row: mean %.1f, max %d, rows > 160: %ld, rows with 0: %ld, overflow rows: %ld of %zu\n",
                    sc / bn, mxc, big, nz, no, bn);
            free(hc); free(ho);
        }
        return knn_fallback_rows(x, norm, fb_list, fb_count, B, C, N, k1, k2, idx64, idx32, st);
    }
    int stride = (int)(tiles * 0.381966f);
    if (stride < 1) stride = 1;
    auto gcd = [](int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; };
    while (gcd(stride, tiles) != 1) ++stride;
    // GCANET_TC_DEBUG=1 (measurement aid, tools/time_knn.py): run the scan pipeline with appends disabled
    // and stop after it -- gives the TMA + MMA + TMEM-read + compare floor of the kernel.
    const char *dbg = GCANET_AID_ENV("GCANET_TC_DEBUG");
    const int dbg_mode = (dbg && dbg[0] >= '1' && dbg[0] <= '3') ? dbg[0] - '0' : 0;
    TcScanArgs sa{norm_pad, Npad, norm, nmax, cand, cand_cnt, overflow, N, k2, tiles, dbg_mode, stride, nullptr, TC_CAP, 0};
    RerankArgs ra{x_nc, norm, nmax, cand, cand_cnt, overflow, idx64, idx32, N, k2, k2 / k1, gcanet_knn_graph_columns(k1, k2),
                  (unordered && k1 == k2) ? 1 : 0, nullptr, TC_CAP, 0, big_list, big_count, fb_list, fb_count};
    if (dbg_mode >= 2 && C == 64)
        rc = dbg_mode == 2 ? launch_tc<64, 2>(tmap_q, tmap_k, sa, ra, B, st) : launch_tc<64, 3>(tmap_q, tmap_k, sa, ra, B, st);
    else
        rc = C == 64 ? launch_tc<64>(tmap_q, tmap_k, sa, ra, B, st) : launch_tc<128>(tmap_q, tmap_k, sa, ra, B, st);
    if (rc) return rc;
    if (sa.debug_no_append) return GCANET_OK;
    return knn_fallback_rows(x, norm, fb_list, fb_count, B, C, N, k1, k2, idx64, idx32, st);
}

}  // namespace gcanet
