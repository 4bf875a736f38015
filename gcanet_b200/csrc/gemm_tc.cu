// fp32-accurate GEMM on the 5th-generation tensor cores for the EdgeConv projections:
//   PQ = X Wcat            [M][K] x [K][N],  K = C  (64),      N = 2 Cout (128 / 256)
//   dX = dPQ Wcat^T        [M][K] x [K][N],  K = 2 Cout,       N = C (64)
// with M = B * N_points (160 000 rows).  C[M][N] = A[M][K] * Bt[N][K]^T, all fp32 in memory.
//
// Both operands are split x = hi + lo (bf16 each) while they are staged into shared memory, and
// hi*hi + hi*lo + lo*hi is accumulated in fp32 TMEM by tcgen05.mma (error ~2^-16 |a||b| per product, the
// level of fp32 summation error at these K).  No bf16 copy of the activations ever exists in HBM:
//   warps 0-7  load fp32 rows of A (coalesced 16-byte loads, two work items ahead), split them, and write the two bf16
//              tiles in the 128-byte-swizzled K-major layout the UMMA descriptors expect (what TMA would have produced)
//   warp  12   issues the MMAs (one elected thread), accumulators double-buffered in TMEM
//   warps 8-11 drain the accumulators: one row per thread, 128-byte segments straight to global memory
// The weight matrix (<= 64 KB as hi|lo bf16) is staged once per CTA; CTAs are persistent over the M tiles.
// The kernel is memory-bound (reads A once, writes C once), which is the point: the CUDA-core sgemm it
// replaces ran at ~33 TFLOP/s fp32 and was compute-bound.
#include "common.cuh"
#include "tc_ptx.cuh"

#include <stdlib.h>

namespace gcanet {

constexpr int GT_BM = 128;            // rows per tile (UMMA M)
constexpr int GT_KB = 64;             // bf16 elements per 128-byte swizzle row = K chunk per pipeline stage
// A stages: as many 32 KB (hi | lo) stages as fit next to the weights, at most 4
__host__ __device__ constexpr int gt_stages(int K, int N) {
    const int left = 200 * 1024 - 2 * (K / 64) * N * 128;
    return left >= 4 * 32768 ? 4 : (left >= 3 * 32768 ? 3 : 2);
}
constexpr int GT_LOADERS = 256;       // loader threads (8 warps)
constexpr int GT_THREADS = GT_LOADERS + 128 + 32;   // 8 loader warps, 4 epilogue warps, 1 MMA warp

__host__ __device__ constexpr uint32_t tmem_cols(int n) { return n <= 32 ? 32 : (n <= 64 ? 64 : (n <= 128 ? 128 : (n <= 256 ? 256 : 512))); }

// byte offset of element (row r, column kc < 64) inside a [rows][64] bf16 tile with the 128-byte swizzle:
// 16-byte chunk index XOR (row mod 8)
__device__ __forceinline__ uint32_t sw128_offset(int r, int kc) {
    return (uint32_t)(r * 128 + ((((kc >> 3) ^ (r & 7)) << 4) | ((kc & 7) << 1)));
}

// four consecutive floats -> four bf16 hi + four bf16 lo, stored as two 8-byte words at (r, kc .. kc+3)
__device__ __forceinline__ void split_store4(uint8_t *hi_tile, uint8_t *lo_tile, int r, int kc, float4 v) {
    const float f[4] = {v.x, v.y, v.z, v.w};
    __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        h[i] = __float2bfloat16_rn(f[i]);
        l[i] = __float2bfloat16_rn(f[i] - __bfloat162float(h[i]));
    }
    const uint32_t off = sw128_offset(r, kc);
    uint2 hw, lw;
    hw.x = (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16);
    hw.y = (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16);
    lw.x = (uint32_t)__bfloat16_as_ushort(l[0]) | ((uint32_t)__bfloat16_as_ushort(l[1]) << 16);
    lw.y = (uint32_t)__bfloat16_as_ushort(l[2]) | ((uint32_t)__bfloat16_as_ushort(l[3]) << 16);
    *reinterpret_cast<uint2 *>(hi_tile + off) = hw;
    *reinterpret_cast<uint2 *>(lo_tile + off) = lw;
}

template <int K, int N>
__global__ void __launch_bounds__(GT_THREADS, 1)
gemm_tc_kernel(const float *__restrict__ A, int lda, const float *__restrict__ Bt, int ldb, float *__restrict__ C, int ldc, int M,
               long long batch_a, long long batch_b, long long batch_c, const float *__restrict__ cbias, int batch_bias,
               int out_bf16) {
    // batched use (blockIdx.y = batch): every batch has its own M x K rows of A, its own weights and, optionally, a row
    // vector cbias[batch][N] added to every output row; element strides between batches are passed in
    A += (size_t)blockIdx.y * batch_a;
    Bt += (size_t)blockIdx.y * batch_b;
    C += (size_t)blockIdx.y * batch_c;
    if (cbias) cbias += (size_t)blockIdx.y * batch_bias;
    static_assert(K % GT_KB == 0 && N % 32 == 0 && N <= 256, "unsupported GEMM shape");
    constexpr int KCH = K / GT_KB;                      // K chunks
    constexpr int GT_STAGES = gt_stages(K, N);
    constexpr int A_TILE = GT_BM * 128;                 // one [128][64] bf16 tile: 16 KB
    constexpr int A_STAGE = 2 * A_TILE;                 // hi + lo
    constexpr int B_BLK = N * 128;                      // one [N][64] bf16 block
    constexpr int B_BYTES = 2 * KCH * B_BLK;            // hi blocks, then lo blocks
    constexpr uint32_t IDESC = umma_idesc_bf16(GT_BM, N);

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *sBm = smem;                                // weights
    uint8_t *sA = smem + B_BYTES;                       // GT_STAGES x (hi | lo)
    uint64_t *bars = reinterpret_cast<uint64_t *>(sA + GT_STAGES * A_STAGE);
    uint64_t *a_full = bars;                            // [STAGES] loaders -> MMA   (128 arrivals)
    uint64_t *a_empty = bars + GT_STAGES;               // [STAGES] MMA -> loaders
    uint64_t *t_full = bars + 2 * GT_STAGES;            // [2] MMA -> epilogue
    uint64_t *t_empty = t_full + 2;                     // [2] epilogue -> MMA       (4 arrivals)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(t_empty + 2);
    // [4 epilogue warps][32 rows][32 floats] staging of the output rows (16-byte chunks XOR-swizzled by the row), 1 KB-aligned
    // behind the barrier block: the accumulator hands every thread ONE row, and thirty-two threads storing 16 bytes each to
    // thirty-two different rows cost the load/store unit 32 cycles per instruction (8.4 tiles x 128 such stores per SM were
    // 17 of the kernel's 39 us); through this buffer a store instruction covers four rows x 128 contiguous bytes
    float *sOut = reinterpret_cast<float *>(sA + GT_STAGES * A_STAGE + 1024);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = (M + GT_BM - 1) / GT_BM;

    if (threadIdx.x == 0) {
        for (int s = 0; s < GT_STAGES; ++s) { mbar_init(&a_full[s], GT_LOADERS); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 4); }
        fence_barrier_init();
    }
    if (warp == GT_THREADS / 32 - 1) tmem_alloc(tmem_slot, tmem_cols(2 * N));
    // weights: Bt[n][k] fp32 -> hi / lo bf16 blocks, swizzled (every thread helps; N*K/4 float4 items)
    for (int e = threadIdx.x; e < N * K / 4; e += GT_THREADS) {
        const int n = e / (K / 4), k4 = (e % (K / 4)) * 4;
        const float4 v = *reinterpret_cast<const float4 *>(Bt + (size_t)n * ldb + k4);
        const int kb = k4 / GT_KB, kc = k4 % GT_KB;
        split_store4(sBm + kb * B_BLK, sBm + (KCH + kb) * B_BLK, n, kc, v);
    }
    fence_proxy_async();                                // generic-proxy writes above are read by the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < GT_LOADERS / 32) {
        // ===================== loaders: fp32 rows -> swizzled bf16 hi / lo tiles =====================
        const int tid = threadIdx.x;                    // 0..255
        const int rsub = tid >> 4, c4 = tid & 15;       // 16 rows x 16 float4 per sweep, 8 sweeps per tile
        int stage = 0;
        uint32_t phase = 0;
        // work items = (tile, K chunk) in order; three register buffers: while item i is split and stored the loads of
        // items i+1 and i+2 are in flight (64 KB per SM -- the kernel is bound by how many bytes it keeps in flight)
        const int my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        const int items = my_tiles * KCH;
        auto issue = [&](int item, float4 (&v)[8]) {
            const int m0 = (blockIdx.x + (item / KCH) * gridDim.x) * GT_BM, kc = item % KCH;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int r = it * 16 + rsub;
                v[it] = m0 + r < M ? __ldg(reinterpret_cast<const float4 *>(A + (size_t)(m0 + r) * lda + kc * GT_KB) + c4)
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        auto consume = [&](const float4 (&v)[8]) {
            mbar_wait(&a_empty[stage], phase ^ 1);
            uint8_t *hi_tile = sA + stage * A_STAGE, *lo_tile = hi_tile + A_TILE;
#pragma unroll
            for (int it = 0; it < 8; ++it) split_store4(hi_tile, lo_tile, it * 16 + rsub, c4 * 4, v[it]);
            fence_proxy_async();
            mbar_arrive(&a_full[stage]);
            if (++stage == GT_STAGES) { stage = 0; phase ^= 1; }
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
    } else if (warp == GT_THREADS / 32 - 1) {
        // ===================== MMA issuer =====================
        // whole warp converged, one elected lane issues (tc_ptx.cuh: elect_one_sync), descriptors advanced from a base
        {
            const uint64_t b_desc0 = make_kmajor_sw128_desc(smem_u32(sBm));
            const uint64_t a_desc0 = make_kmajor_sw128_desc(smem_u32(sA));
            int stage = 0, acc = 0;
            uint32_t phase = 0, accphase = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                mbar_wait(&t_empty[acc], accphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * N;
                for (int kc = 0; kc < KCH; ++kc) {
                    mbar_wait(&a_full[stage], phase);
                    tc_fence_after();
                    if (elect_one_sync()) {
                        const uint64_t a_desc = desc_advance(a_desc0, stage * A_STAGE);
#pragma unroll
                        for (int ks = 0; ks < GT_KB / 16; ++ks) {
                            const uint32_t koff = ks * 32;
                            const uint64_t a_hi = desc_advance(a_desc, koff);
                            const uint64_t a_lo = desc_advance(a_desc, A_TILE + koff);
                            const uint64_t b_hi = desc_advance(b_desc0, kc * B_BLK + koff);
                            const uint64_t b_lo = desc_advance(b_desc0, (KCH + kc) * B_BLK + koff);
                            umma_bf16(d_tmem, a_hi, b_hi, IDESC, (kc | ks) ? 1u : 0u);
                            umma_bf16(d_tmem, a_hi, b_lo, IDESC, 1);
                            umma_bf16(d_tmem, a_lo, b_hi, IDESC, 1);
                        }
                        umma_commit(&a_empty[stage]);
                        // tcgen05.commit covers the MMAs of the EXECUTING thread: the accumulator's commit must come from
                        // the lane that issued them, i.e. from inside the same elected block
                        if (kc == KCH - 1) umma_commit(&t_full[acc]);
                    }
                    __syncwarp();
                    if (++stage == GT_STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; accphase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue: one row per thread =====================
        const int ew = warp & 3;                          // TMEM lane quarter (warps 8..11 -> 0..3)
        int acc = 0;
        uint32_t accphase = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int row = tile * GT_BM + ew * 32 + lane;
            mbar_wait(&t_full[acc], accphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * N;
            __nv_bfloat16 *hrow = reinterpret_cast<__nv_bfloat16 *>(C) + (size_t)row * ldc;     // out_bf16: C is a bf16 matrix
#pragma unroll 1
            for (int ch = 0; ch < N / 32; ++ch) {
                uint32_t v[32];
                tmem_ld32(taddr + ch * 32, v);
                tmem_ld_wait();
                if (row < M && out_bf16) {
                    // rounded to bf16 where it is produced: 32 columns = 64 bytes = four 16-byte stores
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t w[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(v[q * 8 + 2 * e]), __uint_as_float(v[q * 8 + 2 * e + 1]));
                            w[e] = *reinterpret_cast<const uint32_t *>(&h2);
                        }
                        *reinterpret_cast<uint4 *>(hrow + ch * 32 + q * 8) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                } else if (!out_bf16) {
                    // own row -> staging buffer (chunk q of row `lane` at physical chunk q ^ (lane & 7): conflict-free both ways)
                    float4 *st4 = reinterpret_cast<float4 *>(sOut + ew * 1024);
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        st4[lane * 8 + (q ^ (lane & 7))] = make_float4(__uint_as_float(v[q * 4]), __uint_as_float(v[q * 4 + 1]),
                                                                         __uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3]));
                    __syncwarp();
                    // eight lanes per row: one instruction stores four complete 128-byte row segments
                    const int cq = lane & 7;
                    float4 cb = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (cbias) cb = __ldg(reinterpret_cast<const float4 *>(cbias + ch * 32 + cq * 4));
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int r = i * 4 + (lane >> 3);
                        float4 o = st4[r * 8 + (cq ^ (r & 7))];
                        o.x += cb.x; o.y += cb.y; o.z += cb.z; o.w += cb.w;
                        const int grow_ = tile * GT_BM + ew * 32 + r;
                        if (grow_ < M) *reinterpret_cast<float4 *>(C + (size_t)grow_ * ldc + ch * 32 + cq * 4) = o;
                    }
                    __syncwarp();
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[acc]);
            if (++acc == 2) { acc = 0; accphase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == GT_THREADS / 32 - 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols(2 * N));
    }
}

template <int K, int N>
static int launch_gemm_tc(const float *A, int lda, const float *Bt, int ldb, float *C, int ldc, int M, cudaStream_t st,
                          int batches = 1, long long batch_a = 0, long long batch_b = 0, long long batch_c = 0,
                          const float *cbias = nullptr, int batch_bias = 0, int out_bf16 = 0) {
    constexpr int KCH = K / GT_KB;
    constexpr int GT_STAGES = gt_stages(K, N);
    const size_t smem = 1024 + (size_t)2 * KCH * N * 128 + (size_t)GT_STAGES * 2 * GT_BM * 128 + 1024 + 4 * 32 * 32 * sizeof(float);
    auto kern = gemm_tc_kernel<K, N>;
    GCANET_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int ntiles = ceil_div(M, GT_BM);
    int ctas = ntiles < kNumSMs ? ntiles : kNumSMs;
    if (batches > 1) ctas = ctas < ceil_div(2 * kNumSMs, batches) ? ctas : ceil_div(2 * kNumSMs, batches);
    kern<<<dim3(ctas, batches), GT_THREADS, smem, st>>>(A, lda, Bt, ldb, C, ldc, M, batch_a, batch_b, batch_c, cbias, batch_bias, out_bf16);
    GCANET_LAUNCH_OK("gemm_tc_kernel");
    return GCANET_OK;
}

// C[M][N] = A[M][K] Bt[N][K]^T on the tensor cores when the shape is one the EdgeConv layers use
// (K, N in {64, 128, 256}, 16-byte aligned rows); returns +1 (GEMM_TC_NOT_COVERED, never a
// negative gcanet_status) when the shape is not covered and the caller should use the CUDA-core GEMM; a negative return is
// a real error and must be propagated.
// out_bf16 != 0: C is a bf16 matrix [M][ldc] (ldc in elements, a multiple of 8), written rounded to nearest.
int gemm_tc_try(const float *A, int lda, const float *Bt, int ldb, float *C, int ldc, int M, int N, int K, cudaStream_t st,
                int out_bf16) {
    if (M < 1024 || lda % 4 || ldb % 4 || ldc % (out_bf16 ? 8 : 4)) return 1;
    if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(Bt) | reinterpret_cast<uintptr_t>(C)) & 15) return 1;
    if (GCANET_AID_ENV("GCANET_NO_TC_GEMM")) return 1;        // measurement aid
#define GT_CASE(KK, NN) if (K == KK && N == NN) return launch_gemm_tc<KK, NN>(A, lda, Bt, ldb, C, ldc, M, st, 1, 0, 0, 0, nullptr, 0, out_bf16)
    GT_CASE(64, 128);
    GT_CASE(64, 256);
    GT_CASE(128, 128);
    GT_CASE(128, 64);
    GT_CASE(256, 64);
    GT_CASE(256, 128);
    GT_CASE(64, 64);
#undef GT_CASE
    return 1;
}

// Batched form: C[b] = A[b] Bt[b]^T (+ cbias[b] on every row), b < batches; same shapes and return convention.
int gemm_tc_batched_try(const float *A, int lda, long long batch_a, const float *Bt, int ldb, long long batch_b, float *C, int ldc,
                        long long batch_c, const float *cbias, int batch_bias, int M, int N, int K, int batches, cudaStream_t st) {
    if (lda % 4 || ldb % 4 || ldc % 4 || batches < 1 || batches > 65535) return 1;
    if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(Bt) | reinterpret_cast<uintptr_t>(C) |
         reinterpret_cast<uintptr_t>(cbias)) & 15) return 1;
    if ((batch_a | batch_b | batch_c | batch_bias) & 3) return 1;
#define GT_CASE(KK, NN) if (K == KK && N == NN) return launch_gemm_tc<KK, NN>(A, lda, Bt, ldb, C, ldc, M, st, batches, batch_a, batch_b, batch_c, cbias, batch_bias)
    GT_CASE(256, 64);
    GT_CASE(256, 128);
    GT_CASE(128, 64);
    GT_CASE(128, 128);
    GT_CASE(64, 64);
    GT_CASE(64, 128);
#undef GT_CASE
    return 1;
}

// =================================================================================
// Weight gradient: out[64][NY] = X[M][64]^T Y[M][NY], the reduction running over the M = B * N points.
// Split over M, one CTA per split; each CTA accumulates  D[n][m] = sum_p Y[p][n] X[p][m]  (two 128-row accumulators at
// NY = 256) in TMEM with the same bf16 hi/lo split as above, and writes its partial as part[split][m][n].
// Both operands are "MN-major" in memory (the reduction index p is the slow one); instead of MN-major descriptors the
// loaders transpose while they stage: a thread owns one row (n or m) and eight consecutive points, i.e. exactly one
// 16-byte chunk of the swizzled K-major tile, so global loads stay coalesced across the warp (consecutive n) and the
// shared-memory stores are conflict-free (eight consecutive rows hit eight different chunk positions).
// All 16 warps load; thread 0 issues the MMAs of a stage after the block-wide barrier; two stages.
// =================================================================================
constexpr int GTN_THREADS = 512;

__device__ __forceinline__ void split_pack8(const float *v, uint4 &hi, uint4 &lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * q], v[2 * q + 1]);
        const float2 hf = __bfloat1622float2(hh);
        const __nv_bfloat162 ll = __floats2bfloat162_rn(v[2 * q] - hf.x, v[2 * q + 1] - hf.y);
        h[q] = *reinterpret_cast<const uint32_t *>(&hh);
        l[q] = *reinterpret_cast<const uint32_t *>(&ll);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

template <int NY>
__global__ void __launch_bounds__(GTN_THREADS, 1)
gemm_tn_tc_kernel(const float *__restrict__ X, int ldx, const float *__restrict__ Y, int ldy, float *__restrict__ part, int M,
                  int rows_per_split, long long batch_x, long long batch_y) {
    // batched use (blockIdx.y = batch): M rows per batch, partials laid out [batch][split][64][NY]
    X += (size_t)blockIdx.y * batch_x;
    Y += (size_t)blockIdx.y * batch_y;
    part += (size_t)blockIdx.y * gridDim.x * 64 * NY;
    static_assert(NY == 128 || NY == 256, "unsupported width");
    constexpr int TA = NY / 128;                        // accumulator tiles (128 rows of n each)
    constexpr int A_TILE = 128 * 128;                   // [128 n][64 p] bf16
    constexpr int B_TILE = 64 * 128;                    // [64 m][64 p] bf16
    constexpr int STAGE = 2 * TA * A_TILE + 2 * B_TILE; // A hi tiles, A lo tiles, B hi, B lo
    constexpr int UA = NY * 8 / GTN_THREADS;            // (row, chunk) units of Y per thread
    constexpr uint32_t IDESC = umma_idesc_bf16(128, 64);

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 2 * STAGE);
    uint64_t *empty = bars;                             // [2] MMA -> loaders
    uint64_t *done = bars + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 3);

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int start = blockIdx.x * rows_per_split;
    const int end = min(M, start + rows_per_split);
    const int nkb = (end - start + 63) / 64;

    if (t == 0) {
        mbar_init(&empty[0], 1);
        mbar_init(&empty[1], 1);
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 15) tmem_alloc(tmem_slot, tmem_cols(TA * 64));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // this thread's units: Y (n, chunk) x UA, X (m, chunk) x 1
    const int xm = t & 63, xc = t >> 6;
    auto issue = [&](int kb, float (&ya)[UA][8], float (&xa)[8]) {
        const int p0 = start + kb * 64;
#pragma unroll
        for (int i = 0; i < UA; ++i) {
            const int u = t + GTN_THREADS * i, n = u % NY, c = u / NY;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int p = p0 + 8 * c + j;
                ya[i][j] = p < end ? __ldg(Y + (size_t)p * ldy + n) : 0.f;
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int p = p0 + 8 * xc + j;
            xa[j] = p < end ? __ldg(X + (size_t)p * ldx + xm) : 0.f;
        }
    };
    auto store = [&](uint8_t *st, const float (&ya)[UA][8], const float (&xa)[8]) {
        uint4 hi, lo;
#pragma unroll
        for (int i = 0; i < UA; ++i) {
            const int u = t + GTN_THREADS * i, n = u % NY, c = u / NY;
            const int tile = n >> 7, r = n & 127;
            const uint32_t off = (uint32_t)(tile * A_TILE + r * 128 + ((c ^ (r & 7)) << 4));
            split_pack8(ya[i], hi, lo);
            *reinterpret_cast<uint4 *>(st + off) = hi;
            *reinterpret_cast<uint4 *>(st + TA * A_TILE + off) = lo;
        }
        const uint32_t off = (uint32_t)(xm * 128 + ((xc ^ (xm & 7)) << 4));
        split_pack8(xa, hi, lo);
        *reinterpret_cast<uint4 *>(st + 2 * TA * A_TILE + off) = hi;
        *reinterpret_cast<uint4 *>(st + 2 * TA * A_TILE + B_TILE + off) = lo;
    };
    auto mma_stage = [&](uint8_t *st, uint32_t accum) {
        const uint32_t a_addr = smem_u32(st), b_addr = a_addr + 2 * TA * A_TILE;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const uint32_t koff = ks * 32;
            const uint64_t b_hi = make_kmajor_sw128_desc(b_addr + koff);
            const uint64_t b_lo = make_kmajor_sw128_desc(b_addr + B_TILE + koff);
#pragma unroll
            for (int ta = 0; ta < TA; ++ta) {
                const uint64_t a_hi = make_kmajor_sw128_desc(a_addr + ta * A_TILE + koff);
                const uint64_t a_lo = make_kmajor_sw128_desc(a_addr + (TA + ta) * A_TILE + koff);
                const uint32_t d = tmem_base + ta * 64;
                umma_bf16(d, a_hi, b_hi, IDESC, (ks == 0) ? accum : 1u);
                umma_bf16(d, a_hi, b_lo, IDESC, 1);
                umma_bf16(d, a_lo, b_hi, IDESC, 1);
            }
        }
    };

    float ya0[UA][8], xa0[8], ya1[UA][8], xa1[8];
    if (nkb > 0) issue(0, ya0, xa0);
    for (int kb = 0; kb < nkb; kb += 2) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int cur = kb + half;
            if (cur >= nkb) break;
            float (&yc)[UA][8] = half == 0 ? ya0 : ya1;
            float (&xcur)[8] = half == 0 ? xa0 : xa1;
            float (&yn)[UA][8] = half == 0 ? ya1 : ya0;
            float (&xn)[8] = half == 0 ? xa1 : xa0;
            if (cur + 1 < nkb) issue(cur + 1, yn, xn);
            uint8_t *st = smem + half * STAGE;           // stage = cur & 1 = half (kb is even)
            const int use = cur >> 1;                    // how often this stage has been filled before
            if (use > 0) mbar_wait(&empty[half], (uint32_t)((use - 1) & 1));
            store(st, yc, xcur);
            fence_proxy_async();
            __syncthreads();
            if (warp == 0) {
                tc_fence_after();
                if (elect_one_sync()) {
                    mma_stage(st, cur == 0 ? 0u : 1u);
                    umma_commit(&empty[half]);
                    if (cur == nkb - 1) umma_commit(done);      // same lane as the MMAs it has to cover
                }
                __syncwarp();
            }
        }
    }
    if (nkb <= 0 && warp == 0) {                                // nothing was issued: release the waiters
        if (elect_one_sync()) umma_commit(done);
        __syncwarp();
    }
    mbar_wait(done, 0);
    tc_fence_after();

    // epilogue: thread = one output column n, 64 values of m; for each m a warp writes 32 consecutive n
    if (warp < 4 * TA) {
        const int tile = warp >> 2, q = warp & 3;
        const int n = tile * 128 + q * 32 + lane;
        float *po = part + (size_t)blockIdx.x * 64 * NY + n;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + tile * 64;
#pragma unroll 1
        for (int ch = 0; ch < 2; ++ch) {
            uint32_t v[32];
            tmem_ld32(taddr + ch * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int m = 0; m < 32; ++m) po[(size_t)(ch * 32 + m) * NY] = nkb > 0 ? __uint_as_float(v[m]) : 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 15) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols(TA * 64));
    }
}

// part[split][64][N] partials of X[M][64]^T Y[M][N] on the tensor cores; *splits_out = the number of splits written
// (<= max_splits, the caller reduces them).  Returns +1 when the shape is not one this kernel covers.
int gemm_tn_tc_try(const float *X, int ldx, const float *Y, int ldy, float *part, int M, int N, int K, int max_splits,
                   int *splits_out, cudaStream_t st) {
    if (K != 64 || (N != 128 && N != 256) || M < 8192 || max_splits < 1) return 1;
    if (ldx % 4 || ldy % 4 || GCANET_AID_ENV("GCANET_NO_TC_GEMM")) return 1;
    int splits = max_splits < kNumSMs ? max_splits : kNumSMs;
    const int rows = ceil_div(ceil_div(M, splits), 64) * 64;
    splits = ceil_div(M, rows);
    const size_t smem = 1024 + (size_t)2 * (2 * (N / 128) * 128 * 128 + 2 * 64 * 128) + 64;
    if (N == 256) {
        GCANET_CUDA_OK(cudaFuncSetAttribute(gemm_tn_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gemm_tn_tc_kernel<256><<<splits, GTN_THREADS, smem, st>>>(X, ldx, Y, ldy, part, M, rows, 0, 0);
    } else {
        GCANET_CUDA_OK(cudaFuncSetAttribute(gemm_tn_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gemm_tn_tc_kernel<128><<<splits, GTN_THREADS, smem, st>>>(X, ldx, Y, ldy, part, M, rows, 0, 0);
    }
    GCANET_LAUNCH_OK("gemm_tn_tc_kernel");
    *splits_out = splits;
    return GCANET_OK;
}

// Batched Gram-type product: part[b][split][64][256] partials of X[b][M][64]^T Y[b][M][256] with `splits` splits of the
// M rows of every batch (the caller reduces them); rows per split are a multiple of 64.
int gemm_tn_tc_batched(const float *X, int ldx, long long batch_x, const float *Y, int ldy, long long batch_y, float *part, int M,
                       int splits, int batches, cudaStream_t st) {
    if (ldx % 4 || ldy % 4 || splits < 1 || batches < 1 || batches > 65535) return 1;
    const int rows = ceil_div(ceil_div(M, splits), 64) * 64;
    if (ceil_div(M, rows) != splits) return 1;              // the caller sized `part` for exactly this many
    const size_t smem = 1024 + (size_t)2 * (2 * 2 * 128 * 128 + 2 * 64 * 128) + 64;
    GCANET_CUDA_OK(cudaFuncSetAttribute(gemm_tn_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gemm_tn_tc_kernel<256><<<dim3(splits, batches), GTN_THREADS, smem, st>>>(X, ldx, Y, ldy, part, M, rows, batch_x, batch_y);
    GCANET_LAUNCH_OK("gemm_tn_tc_kernel<batched>");
    return GCANET_OK;
}

}  // namespace gcanet
