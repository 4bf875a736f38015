// Feature-space kNN (C = 64 / 128, L2 metric) on the 5th-generation tensor cores.
//
//   prep      x[B][C][N] fp32 -> xs[B][N][2C] bf16 = (hi | lo) with hi = bf16(x), lo = bf16(x - hi),
//             x_nc[B][N][C] fp32 (point-major, for the exact re-rank), |x|^2 (reference order)
//   scan      one CTA per (cloud, 128-query tile), warp-specialised:
//               warp 0    TMA producer: query tile once, then 128-key tiles through an mbarrier ring
//               warp 1    tcgen05.mma issuer: D[128 x 128] (TMEM, fp32) = Qhi Khi^T + Qhi Klo^T + Qlo Khi^T
//                         (bf16 x 3 split, |error| <= ~2^-14 |q||k|), double-buffered accumulators
//               warps 2-5 epilogue: tcgen05.ld one accumulator row per thread, d~ = |k|^2 - 2 q.k,
//                         threshold filter, survivors appended to the row's candidate list;
//                         a full list is compacted by the warp (bisection for an upper bound of the
//                         k-th smallest, keep everything below bound + margin)
//             Every key whose approximate distance is within `margin` (>= 2 x the error bound) of the
//             approximate k-th survives, so the exact k nearest are always among the candidates.
//   rerank    one warp per query: exact fp32 distances of the <= ~k+slack candidates in the
//             reference's expansion arithmetic, rank by (distance, index), write the k best in order.
//   fallback  rows whose candidate list overflowed (massive ties) are redone by the CUDA-core scan.
//
// No N x N matrix and no approximate distance ever decides the result: tensor cores only prune.
#include "common.cuh"

#include <cuda.h>
#include <cuda_bf16.h>
#include <math_constants.h>
#include <stdlib.h>

namespace gcanet {

constexpr unsigned FULLW = 0xffffffffu;
constexpr int TC_BM = 128;            // queries per CTA  (UMMA M)
constexpr int TC_BN = 64;             // keys per tile    (UMMA N); 64 keeps a CTA at ~66 KB so three share an SM
constexpr int TC_KB = 64;             // bf16 elements per 128-byte swizzle row
constexpr int TC_CAP = 256;           // candidate list capacity per query
constexpr int TC_SLACK = 8;           // bisection stops once the bound keeps <= k + slack entries
constexpr int TC_THREADS = 192;       // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue
constexpr int TC_NRING = 8;           // key-norm ring slots (producer is never more than 4 tiles ahead of the epilogue)
constexpr float TC_MARGIN = 7.0e-4f;  // ~2^-10.5 : margin = TC_MARGIN * |q| * max|k|  

// ---------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// same, for the single-thread producer / MMA roles: back off between probes so the spin does
// not steal issue slots from the epilogue warps that share the scheduler
__device__ __forceinline__ void mbar_wait_backoff(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"   // hardware-suspended up to %3 ns
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity), "r"(200000u)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// 1-D bulk copy global -> shared, completion counted on the same mbarrier as the tensor loads
__device__ __forceinline__ void bulk_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate, issued by one thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// K-major operand tile, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO),
// descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B  (cute/arch/mma_sm100_desc.hpp)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                 // LBO (unused for swizzled K-major), canonical value 1
    d |= (uint64_t)(1024 >> 4) << 32;       // SBO
    d |= (uint64_t)1 << 46;                 // version
    d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
    return d;
}
// kind::f16, A = B = bf16, D = fp32, both K-major, M = 128, N = 128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

// ---------------------------------------------------------------------------------
// prep
// ---------------------------------------------------------------------------------
// xs[b][n][0:C] = hi, xs[b][n][C:2C] = lo ; x_nc[b][n][c] = x ; transposes through smem
__global__ void tc_prep_kernel(const float *__restrict__ x, __nv_bfloat16 *__restrict__ xs, float *__restrict__ x_nc,
                               int C, int N) {
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
            size_t row = (size_t)b * N + n;
            xs[row * 2 * C + c] = hi;
            xs[row * 2 * C + C + c] = lo;
            x_nc[row * C + c] = v;
        }
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
        norm_pad[(size_t)b * Npad + n] = v;
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
__device__ __forceinline__ void compact_row(uint2 *ptr, int n, int k, float margin, int lane, int &new_cnt, float &new_thr) {
    constexpr int SL = TC_CAP / 32;
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

template <int C>
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
        if (lane == 0) {
            mbar_wait(a_full, 0);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(sA);
            int stage = 0, acc = 0;
            uint32_t phase = 0, accphase = 0;
            for (int t = 0; t < tiles; ++t) {
                mbar_wait_backoff(&t_empty[acc], accphase ^ 1);
                mbar_wait_backoff(&full[stage], phase);
                tc_fence_after();
                const uint32_t b_addr = smem_u32(sB + stage * TILE_BYTES);
                const uint32_t d_tmem = tmem_base + acc * TC_BN;
                uint32_t accum = 0;
#pragma unroll
                for (int hb = 0; hb < NH; ++hb) {
#pragma unroll
                    for (int ks = 0; ks < TC_KB / 16; ++ks) {
                        const uint32_t koff = ks * 32;        // 16 bf16 = 32 bytes inside the swizzled row
                        const uint64_t a_hi = make_kmajor_sw128_desc(a_addr + hb * ABLK_BYTES + koff);
                        const uint64_t a_lo = make_kmajor_sw128_desc(a_addr + (NH + hb) * ABLK_BYTES + koff);
                        const uint64_t b_hi = make_kmajor_sw128_desc(b_addr + hb * BLK_BYTES + koff);
                        const uint64_t b_lo = make_kmajor_sw128_desc(b_addr + (NH + hb) * BLK_BYTES + koff);
                        umma_bf16(d_tmem, a_hi, b_hi, kIdesc, accum);
                        accum = 1;
                        umma_bf16(d_tmem, a_hi, b_lo, kIdesc, 1);
                        umma_bf16(d_tmem, a_lo, b_hi, kIdesc, 1);
                    }
                }
                umma_commit(&empty[stage]);      // smem slot reusable once these MMAs have read it
                umma_commit(&t_full[acc]);       // accumulator ready for the epilogue
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
        uint2 *buf = a.cand + grow * TC_CAP;
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
            a.cand_cnt[grow] = ovf ? 0 : cnt;
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
};

template <int VEC>
__device__ __forceinline__ void load_row(const float *p, float (&v)[VEC]) {
    if constexpr (VEC == 2) {
        float2 t = __ldg(reinterpret_cast<const float2 *>(p)); v[0] = t.x; v[1] = t.y;
    } else {
        float4 t = __ldg(reinterpret_cast<const float4 *>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
}

template <int C>
__global__ void __launch_bounds__(256) knn_tc_rerank_kernel(RerankArgs a) {
    constexpr int VEC = C / 32;
    constexpr int SL = TC_CAP / 32;
    __shared__ int s_idx[8][TC_CAP];
    __shared__ float s_d[8][TC_CAP];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int q = blockIdx.x * 8 + warp;
    if (q >= a.N) return;
    const size_t grow = (size_t)b * a.N + q;
    if (a.overflow[grow]) return;                         // the fallback kernel writes this row
    const int n = a.cand_cnt[grow];
    const uint2 *cand = a.cand + grow * TC_CAP;
    const float *xb = a.x_nc + (size_t)b * a.N * C;
    const float *nb = a.norm + (size_t)b * a.N;
    int *sl = s_idx[warp];
    float *sd = s_d[warp];

    float qv[VEC];
    load_row<VEC>(xb + (size_t)q * C + lane * VEC, qv);
    const float qn = nb[q];

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
        if (e < n) { uint2 t = cand[e]; ad[s] = __uint_as_float(t.x); aj[s] = (int)t.y; }
    }
    const float margin = TC_MARGIN * sqrtf(qn * a.nmax[b]);
    float keep_below = CUDART_INF_F, sure_below = -CUDART_INF_F;
    if (n > a.k + TC_SLACK) {
        float lo_s;
        keep_below = select_bound<SL>(ad, n, a.k, &lo_s) + margin;
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
        if (keep) sl[m + __popc(msk & ((1u << lane) - 1))] = aj[s];
        if (sure) {
            const size_t o = grow * a.kout + n_sure + __popc(ssk & ((1u << lane) - 1));
            if (a.idx64) a.idx64[o] = aj[s];
            if (a.idx32) a.idx32[o] = aj[s];
        }
        m += __popc(msk);
        n_sure += __popc(ssk);
    }
    const int k_left = a.k - n_sure;                    // >= 1: fewer than k entries lie below lo
    __syncwarp();

    // 2. exact fp32 distances of the m survivors, four at a time (coalesced row loads, one
    //    6-shuffle transposing reduction per four candidates)
    for (int e0 = 0; e0 < m; e0 += 4) {
        float part[4];
        int jj[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            jj[u] = sl[min(e0 + u, m - 1)];
            float xv[VEC];
            load_row<VEC>(xb + (size_t)jj[u] * C + lane * VEC, xv);
            float p = 0.f;
#pragma unroll
            for (int v = 0; v < VEC; ++v) p = fmaf(qv[v], xv[v], p);
            part[u] = p;
        }
        const bool h16 = lane & 16, h8 = lane & 8;
        float k0 = h16 ? part[2] : part[0], k1 = h16 ? part[3] : part[1];
        const float s0 = h16 ? part[0] : part[2], s1 = h16 ? part[1] : part[3];
        k0 += __shfl_xor_sync(FULLW, s0, 16);
        k1 += __shfl_xor_sync(FULLW, s1, 16);
        float kk = h8 ? k1 : k0;
        const float ss = h8 ? k0 : k1;
        kk += __shfl_xor_sync(FULLW, ss, 8);
        kk += __shfl_xor_sync(FULLW, kk, 4);
        kk += __shfl_xor_sync(FULLW, kk, 2);
        kk += __shfl_xor_sync(FULLW, kk, 1);
        // lanes 8u .. 8u+7 now hold the dot product of candidate e0 + u
        const int u = lane >> 3;
        if ((lane & 7) == 0 && e0 + u < m) {
            const int j = u == 0 ? jj[0] : (u == 1 ? jj[1] : (u == 2 ? jj[2] : jj[3]));
            // reference arithmetic: fl(fl(|x_j|^2 - 2 t) + |x_i|^2)
            sd[e0 + u] = __fadd_rn(fmaf(-2.f, kk, nb[j]), qn);
        }
    }
    __syncwarp();

    // 3. rank the survivors by (distance, index) and write the k best in order
    float dv[SL];
    int di[SL], rank[SL];
#pragma unroll
    for (int s = 0; s < SL; ++s) {
        const int e = s * 32 + lane;
        dv[s] = e < m ? sd[e] : CUDART_INF_F;
        di[s] = e < m ? sl[e] : 0x7fffffff;
        rank[s] = 0;
    }
    if (m <= 32) {
        for (int e = 0; e < m; ++e) {
            const float od = sd[e];
            const int oi = sl[e];
            rank[0] += (od < dv[0] || (od == dv[0] && oi < di[0])) ? 1 : 0;
        }
    } else {
        for (int e = 0; e < m; ++e) {
            const float od = sd[e];
            const int oi = sl[e];
#pragma unroll
            for (int s = 0; s < SL; ++s) {
                if (s * 32 >= m) break;
                rank[s] += (od < dv[s] || (od == dv[s] && oi < di[s])) ? 1 : 0;
            }
        }
    }
#pragma unroll
    for (int s = 0; s < SL; ++s) {
        const int e = s * 32 + lane;
        if (e < m && rank[s] < k_left && rank[s] % a.step == 0) {
            const size_t o = grow * a.kout + n_sure + rank[s] / a.step;
            if (a.idx64) a.idx64[o] = di[s];
            if (a.idx32) a.idx32[o] = di[s];
        }
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

size_t knn_tc_workspace_bytes(int B, int C, int N) {
    size_t bn = (size_t)B * N;
    size_t t = 0;
    t += align_up(bn * 2 * C * sizeof(__nv_bfloat16));   // xs
    t += align_up(bn * C * sizeof(float));               // x_nc
    t += align_up(bn * sizeof(float));                   // norm
    t += align_up((size_t)B * (ceil_div(N, TC_BN) * TC_BN) * sizeof(float));   // norm_pad
    t += align_up((size_t)B * sizeof(float));            // nmax
    t += align_up(bn * TC_CAP * sizeof(uint2));          // cand
    t += align_up(bn * sizeof(int));                     // cand_cnt
    t += align_up(bn * sizeof(int));                     // overflow
    return t;
}

bool knn_tc_supported(int C, int N, int k2) {
    return (C == 64 || C == 128) && k2 <= 128 && N >= TC_BM && k2 + TC_SLACK + 64 <= TC_CAP;
}

// declared in knn_select.cu
int launch_sqnorm_public(const float *x, float *out, int B, int C, int Cuse, int N, cudaStream_t st);
int knn_fallback_rows(const float *x, const float *norms, const int *row_filter, int B, int C, int N, int k1, int k2,
                      int64_t *idx64, int32_t *idx32, cudaStream_t st);

template <int C>
static int launch_tc(const CUtensorMap &tmap_q, const CUtensorMap &tmap_k, TcScanArgs sa, RerankArgs ra, int B, cudaStream_t st) {
    constexpr int NBLK = 2 * C / TC_KB;
    constexpr int STAGES = 2;
    const size_t smem = 1024 + (size_t)NBLK * TC_BM * 128 + (size_t)STAGES * NBLK * TC_BN * 128 +
                        TC_NRING * TC_BN * sizeof(float) + 32 * sizeof(uint64_t);
    auto kern = knn_tc_scan_kernel<C>;
    GCANET_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(sa.N, TC_BM), B);
    kern<<<grid, TC_THREADS, smem, st>>>(tmap_q, tmap_k, sa);
    GCANET_LAUNCH_OK("knn_tc_scan_kernel");
    if (sa.debug_no_append) return GCANET_OK;      // measurement aid: scan pipeline only
    dim3 rgrid(ceil_div(sa.N, 8), B);
    knn_tc_rerank_kernel<C><<<rgrid, 256, 0, st>>>(ra);
    GCANET_LAUNCH_OK("knn_tc_rerank_kernel");
    return GCANET_OK;
}

int knn_graph_tensor_cores(const float *x, int B, int C, int N, int k1, int k2, int64_t *idx64, int32_t *idx32,
                           void *ws, int unordered, cudaStream_t st) {
    EncodeTiledFn encode = get_encode_fn();
    if (!encode) { set_error("knn_graph: cuTensorMapEncodeTiled is not available from the driver"); return GCANET_ERR_CUDA; }
    const size_t bn = (size_t)B * N;
    Carver cv(ws);
    __nv_bfloat16 *xs = cv.take<__nv_bfloat16>(bn * 2 * C);
    float *x_nc = cv.take<float>(bn * C);
    float *norm = cv.take<float>(bn);
    const int Npad = ceil_div(N, TC_BN) * TC_BN;
    float *norm_pad = cv.take<float>((size_t)B * Npad);
    float *nmax = cv.take<float>(B);
    uint2 *cand = cv.take<uint2>(bn * TC_CAP);
    int *cand_cnt = cv.take<int>(bn);
    int *overflow = cv.take<int>(bn);

    int rc = launch_sqnorm_public(x, norm, B, C, C, N, st);
    if (rc) return rc;
    {
        dim3 grid(ceil_div(N, 32), ceil_div(C, 32), B), block(32, 8);
        tc_prep_kernel<<<grid, block, 0, st>>>(x, xs, x_nc, C, N);
        GCANET_LAUNCH_OK("tc_prep_kernel");
        tc_normmax_kernel<<<B, 256, 0, st>>>(norm, nmax, norm_pad, N, Npad);
        GCANET_LAUNCH_OK("tc_normmax_kernel");
    }

    // 3-D tensor map over xs: (K = 2C bf16, N rows, B clouds), box = (64, 128, 1), 128-byte swizzle;
    // rows past N are zero-filled, so a partial last tile never reads the next cloud.
    CUtensorMap tmap_q, tmap_k;
    cuuint64_t gdim[3] = {(cuuint64_t)(2 * C), (cuuint64_t)N, (cuuint64_t)B};
    cuuint64_t gstride[2] = {(cuuint64_t)(2 * C) * sizeof(__nv_bfloat16), (cuuint64_t)N * 2 * C * sizeof(__nv_bfloat16)};
    cuuint32_t estride[3] = {1, 1, 1};
    cuuint32_t box_q[3] = {(cuuint32_t)TC_KB, (cuuint32_t)TC_BM, 1};
    cuuint32_t box_k[3] = {(cuuint32_t)TC_KB, (cuuint32_t)TC_BN, 1};
    CUresult cr = encode(&tmap_q, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, xs, gdim, gstride, box_q, estride,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr == CUDA_SUCCESS)
        cr = encode(&tmap_k, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, xs, gdim, gstride, box_k, estride,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { set_error("knn_graph: cuTensorMapEncodeTiled failed (%d)", (int)cr); return GCANET_ERR_CUDA; }

    const int tiles = ceil_div(N, TC_BN);
    int stride = (int)(tiles * 0.381966f);
    if (stride < 1) stride = 1;
    auto gcd = [](int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; };
    while (gcd(stride, tiles) != 1) ++stride;
    // GCANET_TC_DEBUG=1 (measurement aid, tools/time_knn.py): run the scan pipeline with appends disabled
    // and stop after it -- gives the TMA + MMA + TMEM-read + compare floor of the kernel.
    const char *dbg = getenv("GCANET_TC_DEBUG");
    TcScanArgs sa{norm_pad, Npad, norm, nmax, cand, cand_cnt, overflow, N, k2, tiles, (dbg && dbg[0] == '1') ? 1 : 0, stride};
    RerankArgs ra{x_nc, norm, nmax, cand, cand_cnt, overflow, idx64, idx32, N, k2, k2 / k1, gcanet_knn_graph_columns(k1, k2),
                  (unordered && k1 == k2) ? 1 : 0};
    rc = C == 64 ? launch_tc<64>(tmap_q, tmap_k, sa, ra, B, st) : launch_tc<128>(tmap_q, tmap_k, sa, ra, B, st);
    if (rc) return rc;
    if (sa.debug_no_append) return GCANET_OK;
    return knn_fallback_rows(x, norm, overflow, B, C, N, k1, k2, idx64, idx32, st);
}

}  // namespace gcanet
