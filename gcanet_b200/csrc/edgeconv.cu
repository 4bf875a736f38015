// Fused EdgeConv block: graph feature -> 1x1 conv -> GroupNorm -> LeakyReLU -> max over k,
// forward and backward, without the [B][2C][N][k] edge tensor or the [B][Cout][N][k]
// activation ever existing in memory.
//
// Identities (DESIGN.md section 3 derives them; tests check them against autograd):
//   (1) W [x_j - x_i ; x_i] = W1 x_j + (W2 - W1) x_i  =: P_j + Q_i,  so the per-edge conv
//       output is y_ik = P[idx(i,k)] + Q[i] with [P|Q] = X Wcat one small GEMM per layer.
//   (2) GroupNorm is a per-(sample,channel) affine map a*y + b with sign(a) = sign(gamma) and
//       LeakyReLU is increasing, so  max_k LReLU(GN(y_ik)) = LReLU(GN(a >= 0 ? max_k y : min_k y)).
//       One gather pass therefore yields everything: max, min (+ their k), sum_k y, and the
//       group moments sum y, sum y^2 for the GroupNorm statistics.
//   (3) backward: dy_ik = [k = k*] rstd gamma du + A_g + K_g y_ik  (the last two terms are the
//       GroupNorm mean/variance paths, dense over all edges but affine in y), hence
//         dQ_i = rstd gamma du_i + k A_g + K_g sum_k y_ik
//         dP_j = sum_{(i,k)->j} ([k = k*] rstd gamma du_i + A_g + K_g Q_i) + deg_j K_g P_j
//       i.e. one scatter pass over the edges (vector atomics into an L2-resident buffer), then
//       dX = [dP|dQ] Wcat^T and dWcat = X^T [dP|dQ] as two small GEMMs.
//
// Rooflines: the gather/scatter passes move B*N*k rows of Cout*4 bytes through L2 (the cloud's
// P matrix is N*Cout*4 bytes, L2-resident) and B*N*Cout*~14 bytes through HBM; the GEMMs are
// 2*B*N*C*2Cout FLOP, k times fewer than the per-edge formulation.
#include "common.cuh"

#include <type_traits>

#include <cuda_bf16.h>

#include <cstdlib>
#include <math_constants.h>

namespace gcanet {

// gemm_tc.cu: C[M][N] = A[M][K] Bt[N][K]^T on the tensor cores; -1 = shape not covered, use the CUDA-core GEMM
int gemm_tc_try(const float *A, int lda, const float *Bt, int ldb, float *C, int ldc, int M, int N, int K, cudaStream_t st,
                int out_bf16 = 0);
int gemm_tn_tc_try(const float *X, int ldx, const float *Y, int ldy, float *part, int M, int N, int K, int max_splits, int *splits_out,
                   cudaStream_t st);

constexpr unsigned FULLM = 0xffffffffu;
constexpr int kGWarps = 8;           // warps per CTA in the per-point kernels
constexpr int kPtsPerWarp = 4;       // points each warp handles
constexpr int kPtsPerCta = kGWarps * kPtsPerWarp;

// ---------------------------------------------------------------------------------
// weights:  Wcat[c][o] = W[o][c],  Wcat[c][Cout+o] = W[o][C+c] - W[o][c]   (rows c >= C zero)
//           WcatT[n][c] = Wcat[c][n]
// ---------------------------------------------------------------------------------
__global__ void prep_wcat_kernel(const float *__restrict__ W, float *__restrict__ wcat, float *__restrict__ wcatT,
                                 int C, int ldx, int Cout) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int total = ldx * 2 * Cout;
    if (t >= total) return;
    int c = t / (2 * Cout), n = t % (2 * Cout);
    float v = 0.f;
    if (c < C) {
        if (n < Cout) v = W[(size_t)n * 2 * C + c];
        else { int o = n - Cout; v = W[(size_t)o * 2 * C + C + c] - W[(size_t)o * 2 * C + c]; }
    }
    if (wcat) wcat[t] = v;
    if (wcatT) wcatT[(size_t)n * ldx + c] = v;
}

// dW[o][c] = dWcat[c][o] - dWcat[c][Cout+o];  dW[o][C+c] = dWcat[c][Cout+o]
__global__ void unprep_dw_kernel(const float *__restrict__ dwcat, float *__restrict__ dW, int C, int Cout) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Cout * C) return;
    int o = t / C, c = t % C;
    float dp = dwcat[(size_t)c * 2 * Cout + o], dq = dwcat[(size_t)c * 2 * Cout + Cout + o];
    dW[(size_t)o * 2 * C + c] = dp - dq;
    dW[(size_t)o * 2 * C + C + c] = dq;
}

// ---------------------------------------------------------------------------------
// fp32 GEMM on CUDA cores.   TA = false: C[M][N] = A[M][K] B[K][N]
//                            TA = true : C[K][N] (+= over M-splits) = A[M][K]^T B[M][N]
// Tile 128 x 64 x 16, 256 threads, 8 x 4 outputs per thread.  K, N, lda, ldb, ldc % 4 == 0.
// ---------------------------------------------------------------------------------
constexpr int BM = 128, BN = 64, BK = 16;

template <bool TA>
__global__ void __launch_bounds__(256) sgemm_kernel(const float *__restrict__ A, const float *__restrict__ Bm,
                                                    float *__restrict__ Cm, int M, int N, int K, int lda, int ldb,
                                                    int ldc, int rows_per_split) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int n0 = blockIdx.x * BN;
    const int r0 = blockIdx.y * BM;            // output-row tile (M for NN, K for TN)
    int red_lo = 0, red_hi = TA ? M : K;       // reduction range
    if (TA) { red_lo = blockIdx.z * rows_per_split; red_hi = min(M, red_lo + rows_per_split); }

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = red_lo; k0 < red_hi; k0 += BK) {
        if (!TA) {
            // A tile: rows r0..r0+127, cols k0..k0+15 -> As[k][row]
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                int f = tid + i * 256;
                int row = f >> 2, kq = f & 3;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r0 + row < M && k0 + kq * 4 < red_hi)
                    v = *reinterpret_cast<const float4 *>(A + (size_t)(r0 + row) * lda + k0 + kq * 4);
                As[kq * 4 + 0][row] = v.x; As[kq * 4 + 1][row] = v.y;
                As[kq * 4 + 2][row] = v.z; As[kq * 4 + 3][row] = v.w;
            }
        } else {
            // A tile: reduction rows k0..k0+15, output rows (columns of A) r0..r0+127 -> As[k][row]
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                int f = tid + i * 256;
                int kr = f >> 5, cq = f & 31;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k0 + kr < red_hi && r0 + cq * 4 < K)
                    v = *reinterpret_cast<const float4 *>(A + (size_t)(k0 + kr) * lda + r0 + cq * 4);
                *reinterpret_cast<float4 *>(&As[kr][cq * 4]) = v;
            }
        }
        {
            int kr = tid >> 4, nq = tid & 15;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k0 + kr < red_hi && n0 + nq * 4 < N)
                v = *reinterpret_cast<const float4 *>(Bm + (size_t)(k0 + kr) * ldb + n0 + nq * 4);
            *reinterpret_cast<float4 *>(&Bs[kr][nq * 4]) = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float4 a0 = *reinterpret_cast<const float4 *>(&As[kk][ty * 8]);
            float4 a1 = *reinterpret_cast<const float4 *>(&As[kk][ty * 8 + 4]);
            float4 b4 = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
            float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
    const int out_rows = TA ? K : M;
    float *Cbase = Cm + (TA ? (size_t)blockIdx.z * K * ldc : 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int r = r0 + ty * 8 + i;
        int c = n0 + tx * 4;
        if (r < out_rows && c < N)
            *reinterpret_cast<float4 *>(Cbase + (size_t)r * ldc + c) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
}

// sums `splits` partial [rows][ld] matrices
// (eight warps per 32 outputs, each a strided eighth of the splits with coalesced loads, combined in a fixed order
// through shared memory: the sum is deterministic and no thread walks all `splits` partials serially)
__global__ void __launch_bounds__(256) reduce_splits_kernel(const float *__restrict__ part, float *__restrict__ out, int count, int splits) {
    __shared__ double red[8][32];
    const int sub = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * 32 + lane;
    double s = 0.0;
    if (t < count)
        for (int z = sub; z < splits; z += 8) s += (double)part[(size_t)z * count + t];
    red[sub][lane] = s;
    __syncthreads();
    if (sub == 0 && t < count) {
        double r = red[0][lane];
#pragma unroll
        for (int w = 1; w < 8; ++w) r += red[w][lane];
        out[t] = (float)r;
    }
}

static int launch_sgemm_nn(const float *A, const float *Bm, float *Cm, int M, int N, int K, int lda, int ldb, int ldc,
                           cudaStream_t st) {
    dim3 grid(ceil_div(N, BN), ceil_div(M, BM), 1);
    sgemm_kernel<false><<<grid, 256, 0, st>>>(A, Bm, Cm, M, N, K, lda, ldb, ldc, 0);
    GCANET_LAUNCH_OK("sgemm_kernel<NN>");
    return GCANET_OK;
}

// ---------------------------------------------------------------------------------
// out[K][N] = A[M][K]^T B[M][N]  (weight gradient: reduction over the B*N points).
// Split over M; every split writes its own [K][N] partial, reduced afterwards in a fixed order.
//   K <= 8 : each thread owns one output column and K accumulators, rows are streamed
//   else   : 64 x 64 output tiles, 16 rows of M per step, 4 x 4 outputs per thread
// ---------------------------------------------------------------------------------
constexpr int TN_T = 64, TN_BK = 16;

__global__ void __launch_bounds__(256) sgemm_tn64_kernel(const float *__restrict__ A, const float *__restrict__ Bm,
                                                         float *__restrict__ part, int M, int N, int K, int lda,
                                                         int ldb, int rows_per_split) {
    __shared__ __align__(16) float As[TN_BK][TN_T];
    __shared__ __align__(16) float Bs[TN_BK][TN_T];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int n0 = blockIdx.x * TN_T, c0 = blockIdx.y * TN_T;
    const int lo = blockIdx.z * rows_per_split, hi = min(M, lo + rows_per_split);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int kr = tid >> 4, q4 = (tid & 15) * 4;       // one float4 of A and one of B per thread per step
    for (int m0 = lo; m0 < hi; m0 += TN_BK) {
        float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
        if (m0 + kr < hi) {
            if (c0 + q4 < K) va = *reinterpret_cast<const float4 *>(A + (size_t)(m0 + kr) * lda + c0 + q4);
            if (n0 + q4 < N) vb = *reinterpret_cast<const float4 *>(Bm + (size_t)(m0 + kr) * ldb + n0 + q4);
        }
        *reinterpret_cast<float4 *>(&As[kr][q4]) = va;
        *reinterpret_cast<float4 *>(&Bs[kr][q4]) = vb;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TN_BK; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
    float *out = part + (size_t)blockIdx.z * K * N;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = c0 + ty * 4 + i, c = n0 + tx * 4;
        if (r < K && c < N) *reinterpret_cast<float4 *>(out + (size_t)r * N + c) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
}

// K == 64 and N a multiple of 128 (the feature layers): one CTA owns all 64 rows and NT = 128 / 256 columns of the
// output, 8 x (NT / 32) accumulators per thread (16 FMA per shared-memory load instead of 8), the next 16 rows of
// A and B are fetched into registers while the current ones are multiplied, and A is read once per NT columns.
template <int NT>
__global__ void __launch_bounds__(256) sgemm_tn_wide_kernel(const float *__restrict__ A, const float *__restrict__ Bm,
                                                            float *__restrict__ part, int M, int N, int lda, int ldb,
                                                            int rows_per_split) {
    constexpr int CW = NT / 32;                          // columns per thread: 4 or 8 (one or two float4)
    constexpr int BV = NT / 64;                          // float4 of B per thread per step
    __shared__ __align__(16) float As[TN_BK][64];
    __shared__ __align__(16) float Bs[TN_BK][NT];
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int n0 = blockIdx.x * NT;
    const int lo = blockIdx.z * rows_per_split, hi = min(M, lo + rows_per_split);
    float acc[8][CW];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < CW; ++j) acc[i][j] = 0.f;
    const int ar = tid >> 4, aq = (tid & 15) * 4;        // A: 16 rows x 16 float4
    float4 va, vb[BV];
    auto fetch = [&](int m0) {
        va = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m0 + ar < hi) va = __ldg(reinterpret_cast<const float4 *>(A + (size_t)(m0 + ar) * lda + aq));
#pragma unroll
        for (int u = 0; u < BV; ++u) {
            const int e = tid + u * 256;                 // float4 index inside the 16 x NT block
            const int br = e / (NT / 4), bq = (e % (NT / 4)) * 4;
            vb[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m0 + br < hi) vb[u] = __ldg(reinterpret_cast<const float4 *>(Bm + (size_t)(m0 + br) * ldb + n0 + bq));
        }
    };
    fetch(lo);
    for (int m0 = lo; m0 < hi; m0 += TN_BK) {
        *reinterpret_cast<float4 *>(&As[ar][aq]) = va;
#pragma unroll
        for (int u = 0; u < BV; ++u) {
            const int e = tid + u * 256;
            *reinterpret_cast<float4 *>(&Bs[e / (NT / 4)][(e % (NT / 4)) * 4]) = vb[u];
        }
        __syncthreads();
        if (m0 + TN_BK < hi) fetch(m0 + TN_BK);
#pragma unroll
        for (int kk = 0; kk < TN_BK; ++kk) {
            float a[8], bb[CW];
            *reinterpret_cast<float4 *>(a) = *reinterpret_cast<const float4 *>(&As[kk][ty * 8]);
            *reinterpret_cast<float4 *>(a + 4) = *reinterpret_cast<const float4 *>(&As[kk][ty * 8 + 4]);
#pragma unroll
            for (int h = 0; h < CW / 4; ++h)
                *reinterpret_cast<float4 *>(bb + 4 * h) = *reinterpret_cast<const float4 *>(&Bs[kk][h * 128 + tx * 4]);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < CW; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
    float *out = part + (size_t)blockIdx.z * 64 * N;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int h = 0; h < CW / 4; ++h)
            *reinterpret_cast<float4 *>(out + (size_t)(ty * 8 + i) * N + n0 + h * 128 + tx * 4) =
                make_float4(acc[i][4 * h], acc[i][4 * h + 1], acc[i][4 * h + 2], acc[i][4 * h + 3]);
}

// K <= 8.  blockDim = 256; thread t owns column n = t % N2 of row group t / N2 (N2 = N rounded to the block).
__global__ void __launch_bounds__(256) sgemm_tn_smallk_kernel(const float *__restrict__ A, const float *__restrict__ Bm,
                                                              float *__restrict__ part, int M, int N, int K, int lda,
                                                              int ldb, int rows_per_split) {
    __shared__ float red[8][256];
    const int groups = 256 / N > 0 ? 256 / N : 1;       // row groups working on interleaved rows
    const int col = threadIdx.x % N, grp = threadIdx.x / N;
    const bool live = grp < groups && (N <= 256);
    const int lo = blockIdx.x * rows_per_split, hi = min(M, lo + rows_per_split);
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;
    if (live) {
        for (int m = lo + grp; m < hi; m += groups) {
            const float bv = Bm[(size_t)m * ldb + col];
            const float *ar = A + (size_t)m * lda;
#pragma unroll
            for (int c = 0; c < 8; ++c)
                if (c < K) acc[c] = fmaf(ar[c], bv, acc[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) red[c][threadIdx.x] = live ? acc[c] : 0.f;
    __syncthreads();
    for (int e = threadIdx.x; e < K * N; e += 256) {
        const int c = e / N, n = e % N;
        float s = 0.f;
        for (int g = 0; g < groups; ++g) s += red[c][g * N + n];
        part[(size_t)blockIdx.x * K * N + e] = s;
    }
}

static int tn_wide_nt(int N, int K) { return K == 64 ? (N % 256 == 0 ? 256 : (N % 128 == 0 ? 128 : 0)) : 0; }

static int tn_splits(int M, int N, int K) {
    int splits;
    if (K <= 8 && N <= 256) splits = 4 * kNumSMs;
    else if (tn_wide_nt(N, K)) splits = (2 * kNumSMs + N / tn_wide_nt(N, K) - 1) / (N / tn_wide_nt(N, K));
    else {
        int tiles = ceil_div(N, TN_T) * ceil_div(K, TN_T);
        splits = (4 * kNumSMs + tiles - 1) / tiles;
    }
    int max_splits = ceil_div(M, 4 * TN_BK);
    if (splits > max_splits) splits = max_splits;
    return splits < 1 ? 1 : splits;
}

// out[K][N] = A[M][K]^T B[M][N]; part must hold tn_splits * K * N floats
static int launch_sgemm_tn(const float *A, const float *Bm, float *out, float *part, int M, int N, int K, int lda,
                           int ldb, cudaStream_t st) {
    int splits = tn_splits(M, N, K);
    // feature layers: tensor cores (bf16x3 split, fp32 accumulate)
    int tc_splits = 0;
    const int rc_tc = gemm_tn_tc_try(A, lda, Bm, ldb, part, M, N, K, splits, &tc_splits, st);
    if (rc_tc < 0) return rc_tc;
    if (rc_tc == GCANET_OK) {
        reduce_splits_kernel<<<ceil_div(K * N, 32), 256, 0, st>>>(part, out, K * N, tc_splits);
        GCANET_LAUNCH_OK("reduce_splits_kernel");
        return GCANET_OK;
    }
    int rows = ceil_div(ceil_div(M, splits), TN_BK) * TN_BK;
    splits = ceil_div(M, rows);
    if (K <= 8 && N <= 256) {
        sgemm_tn_smallk_kernel<<<splits, 256, 0, st>>>(A, Bm, part, M, N, K, lda, ldb, rows);
        GCANET_LAUNCH_OK("sgemm_tn_smallk_kernel");
    } else if (tn_wide_nt(N, K) == 256) {
        sgemm_tn_wide_kernel<256><<<dim3(N / 256, 1, splits), 256, 0, st>>>(A, Bm, part, M, N, lda, ldb, rows);
        GCANET_LAUNCH_OK("sgemm_tn_wide_kernel");
    } else if (tn_wide_nt(N, K) == 128) {
        sgemm_tn_wide_kernel<128><<<dim3(N / 128, 1, splits), 256, 0, st>>>(A, Bm, part, M, N, lda, ldb, rows);
        GCANET_LAUNCH_OK("sgemm_tn_wide_kernel");
    } else {
        dim3 grid(ceil_div(N, TN_T), ceil_div(K, TN_T), splits);
        sgemm_tn64_kernel<<<grid, 256, 0, st>>>(A, Bm, part, M, N, K, lda, ldb, rows);
        GCANET_LAUNCH_OK("sgemm_tn64_kernel");
    }
    int count = K * N;
    reduce_splits_kernel<<<ceil_div(count, 32), 256, 0, st>>>(part, out, count, splits);
    GCANET_LAUNCH_OK("reduce_splits_kernel");
    return GCANET_OK;
}

// ---------------------------------------------------------------------------------
// forward gather pass
// ---------------------------------------------------------------------------------
template <int VEC>
struct VecIO;
template <>
struct VecIO<1> {
    static __device__ __forceinline__ void ld(const float *p, float *v) { v[0] = __ldg(p); }
    static __device__ __forceinline__ void st(float *p, const float *v) { p[0] = v[0]; }
    static __device__ __forceinline__ void red(float *p, const float *v) { atomicAdd(p, v[0]); }
};
template <>
struct VecIO<2> {
    static __device__ __forceinline__ void ld(const float *p, float *v) {
        float2 t = __ldg(reinterpret_cast<const float2 *>(p)); v[0] = t.x; v[1] = t.y;
    }
    static __device__ __forceinline__ void st(float *p, const float *v) { *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]); }
    static __device__ __forceinline__ void red(float *p, const float *v) { atomicAdd(reinterpret_cast<float2 *>(p), make_float2(v[0], v[1])); }
};
template <>
struct VecIO<4> {
    static __device__ __forceinline__ void ld(const float *p, float *v) {
        float4 t = __ldg(reinterpret_cast<const float4 *>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void st(float *p, const float *v) { *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]); }
    static __device__ __forceinline__ void red(float *p, const float *v) { atomicAdd(reinterpret_cast<float4 *>(p), make_float4(v[0], v[1], v[2], v[3])); }
};
template <>
struct VecIO<8> {
    static __device__ __forceinline__ void ld(const float *p, float *v) { VecIO<4>::ld(p, v); VecIO<4>::ld(p + 4, v + 4); }
    static __device__ __forceinline__ void st(float *p, const float *v) { VecIO<4>::st(p, v); VecIO<4>::st(p + 4, v + 4); }
    static __device__ __forceinline__ void red(float *p, const float *v) { VecIO<4>::red(p, v); VecIO<4>::red(p + 4, v + 4); }
};

// VEC consecutive bf16 -> fp32 (exact: a shift)
template <int VEC>
__device__ __forceinline__ void ld_bf16(const __nv_bfloat16 *p, float *v) {
    if constexpr (VEC == 1) {
        v[0] = __bfloat162float(*p);
    } else if constexpr (VEC == 2) {
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(p));
        v[0] = __uint_as_float(w << 16); v[1] = __uint_as_float(w & 0xffff0000u);
    } else if constexpr (VEC == 4) {
        const uint2 w = __ldg(reinterpret_cast<const uint2 *>(p));
        v[0] = __uint_as_float(w.x << 16); v[1] = __uint_as_float(w.x & 0xffff0000u);
        v[2] = __uint_as_float(w.y << 16); v[3] = __uint_as_float(w.y & 0xffff0000u);
    } else {
        const uint4 w = __ldg(reinterpret_cast<const uint4 *>(p));
        v[0] = __uint_as_float(w.x << 16); v[1] = __uint_as_float(w.x & 0xffff0000u);
        v[2] = __uint_as_float(w.y << 16); v[3] = __uint_as_float(w.y & 0xffff0000u);
        v[4] = __uint_as_float(w.z << 16); v[5] = __uint_as_float(w.z & 0xffff0000u);
        v[6] = __uint_as_float(w.w << 16); v[7] = __uint_as_float(w.w & 0xffff0000u);
    }
}
template <int VEC, bool BF16>
__device__ __forceinline__ void ld_pq(const void *base, size_t elem, float *v) {
    if constexpr (BF16) ld_bf16<VEC>(reinterpret_cast<const __nv_bfloat16 *>(base) + elem, v);
    else VecIO<VEC>::ld(reinterpret_cast<const float *>(base) + elem, v);
}
// row j of a matrix whose rows are `stride_bytes` apart, from a byte pointer to row 0: one IMAD.WIDE.U32 per address
// (a typed pointer plus a 32-bit element offset costs five instructions: multiply, add, carry, shift, high word)
template <int VEC, bool BF16>
__device__ __forceinline__ void ld_pq_row(const char *row0, unsigned j, unsigned stride_bytes, float *v) {
    const char *p = row0 + (unsigned long long)j * stride_bytes;
    if constexpr (BF16) ld_bf16<VEC>(reinterpret_cast<const __nv_bfloat16 *>(p), v);
    else VecIO<VEC>::ld(reinterpret_cast<const float *>(p), v);
}

struct FwdArgs {
    const void *pq;        // [B][N][2*Cout] fp32, or bf16 in the bf16-storage mode
    const int32_t *idx;    // [B][N][k]
    const float *gamma;    // [Cout]  only its sign is used: the extreme that survives max_k(LReLU(GN(.)))
    float *ysel, *ysum;    // [B][N][Cout]  selected pre-norm value (max_k y if gamma >= 0 else min_k y), sum_k y
    unsigned char *arg;    // [B][N][Cout]  the k that attains ysel
    double *part;          // [B][nblk][G][2]
    int N, Cout, k, G;
};

// One edge of the gather: z = sg p + q, running max with its slot, sum and sum of squares.  The three arithmetic steps
// use the packed fp32x2 instructions of sm_100 (FFMA2 / FADD2: two IEEE-rounded results per issue slot, bit-identical to
// the scalar forms) -- the kernel is issue-bound (72 % issue-active), and this takes 6 of its 33 instructions per edge
// at four channels per lane.
template <int VEC>
__device__ __forceinline__ void edge_accumulate(const float *p, const float *sg, const float *q, float *zmax, int *kbest,
                                                float *vsum, float *vsq, int slot) {
    if constexpr (VEC % 2 == 0) {
#pragma unroll
        for (int h = 0; h < VEC; h += 2) {
            const float2 z = __ffma2_rn(make_float2(sg[h], sg[h + 1]), make_float2(p[h], p[h + 1]), make_float2(q[h], q[h + 1]));
            const float2 sm = __fadd2_rn(make_float2(vsum[h], vsum[h + 1]), z);
            const float2 sq = __ffma2_rn(z, z, make_float2(vsq[h], vsq[h + 1]));
            vsum[h] = sm.x; vsum[h + 1] = sm.y;
            vsq[h] = sq.x; vsq[h + 1] = sq.y;
            if (z.x > zmax[h]) { zmax[h] = z.x; kbest[h] = slot; }
            if (z.y > zmax[h + 1]) { zmax[h + 1] = z.y; kbest[h + 1] = slot; }
        }
    } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const float z = fmaf(sg[v], p[v], q[v]);
            if (z > zmax[v]) { zmax[v] = z; kbest[v] = slot; }
            vsum[v] += z;
            vsq[v] = fmaf(z, z, vsq[v]);
        }
    }
}

template <int I, int N, typename F>
__device__ __forceinline__ void gather_static_for(F &&f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        gather_static_for<I + 1, N>(f);
    }
}

template <int VEC, bool BF16 = false>
__global__ void __launch_bounds__(kGWarps * 32) edge_gather_reduce_kernel(FwdArgs a) {
    __shared__ double red[kGWarps * 32][2];
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Cout = a.Cout, k = a.k;
    const int c0 = lane * VEC;
    // rows are addressed with 32-bit element offsets inside the cloud's [N][2 Cout] matrix (N * 2 Cout < 2^30, checked
    // by check_desc): two instructions per edge instead of a seven-instruction 64-bit multiply
    const size_t cloud0 = (size_t)b * a.N * 2 * Cout;
    constexpr unsigned ELEM = BF16 ? 2u : 4u;
    const char *row0 = reinterpret_cast<const char *>(a.pq) + (cloud0 + c0) * ELEM;      // channel c0 of the cloud's first row
    asm volatile("" : "+l"(row0));        // one 64-bit register pair: keeps the compiler from re-adding the kernel argument per edge
    const unsigned stride_b = 2u * (unsigned)Cout * ELEM;
    double s1 = 0.0, s2 = 0.0;
    float sg[VEC];
    VecIO<VEC>::ld(a.gamma + c0, sg);
#pragma unroll
    for (int v = 0; v < VEC; ++v) sg[v] = sg[v] < 0.f ? -1.f : 1.f;

    for (int pi = 0; pi < kPtsPerWarp; ++pi) {
        const int i = blockIdx.x * kPtsPerCta + warp * kPtsPerWarp + pi;
        if (i >= a.N) break;
        // z = sign(gamma) * y: one running max covers both signs (y itself is exact: only a sign flip)
        // with sq = sg * q:  z = fma(sg, p, sq) = sg * fl(p + q) exactly (sg = +-1, rounding is symmetric), sum_k y =
        // sg * sum_k z and y^2 = z^2 -- one instruction per edge and channel less than forming y first
        float q[VEC], zmax[VEC], vsum[VEC], vsq[VEC];
        int kbest[VEC];
        ld_pq<VEC, BF16>(a.pq, cloud0 + (size_t)i * 2 * Cout + Cout + c0, q);
#pragma unroll
        for (int v = 0; v < VEC; ++v) { zmax[v] = -CUDART_INF_F; vsum[v] = 0.f; vsq[v] = 0.f; kbest[v] = 0; q[v] *= sg[v]; }
        const int32_t *ip = a.idx + ((size_t)b * a.N + i) * k;
        int base = 0;
        if (k >= 32) {
            // the first 32 edges with compile-time lane and slot numbers: the shuffle's source lane and the slot written
            // into kbest are immediates (one VIADD each per edge in the generic loop below; the kernel is issue-bound)
            const int myj = ip[lane];
            gather_static_for<0, 4>([&](auto tb) {
                constexpr int t = decltype(tb)::value * 8;
                float p[8][VEC];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    int j = __shfl_sync(FULLM, myj, t + u);
                    ld_pq_row<VEC, BF16>(row0, (unsigned)j, stride_b, p[u]);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) edge_accumulate<VEC>(p[u], sg, q, zmax, kbest, vsum, vsq, t + u);
            });
            base = 32;
        }
        for (; base < k; base += 32) {
            const int cnt = min(32, k - base);
            const int myj = lane < cnt ? ip[base + lane] : 0;
            int t = 0;
            for (; t + 8 <= cnt; t += 8) {
                float p[8][VEC];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    int j = __shfl_sync(FULLM, myj, t + u);
                    ld_pq_row<VEC, BF16>(row0, (unsigned)j, stride_b, p[u]);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) edge_accumulate<VEC>(p[u], sg, q, zmax, kbest, vsum, vsq, base + t + u);
            }
            for (; t < cnt; ++t) {
                float p[VEC];
                int j = __shfl_sync(FULLM, myj, t);
                ld_pq_row<VEC, BF16>(row0, (unsigned)j, stride_b, p);
                edge_accumulate<VEC>(p, sg, q, zmax, kbest, vsum, vsq, base + t);
            }
        }
        const size_t o = ((size_t)b * a.N + i) * Cout + c0;
        float ys[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) { ys[v] = sg[v] * zmax[v]; vsum[v] *= sg[v]; }     // back from z to y
        VecIO<VEC>::st(a.ysel + o, ys);
        VecIO<VEC>::st(a.ysum + o, vsum);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            a.arg[o + v] = (unsigned char)kbest[v];
            s1 += (double)vsum[v];
            s2 += (double)vsq[v];
        }
    }
    red[threadIdx.x][0] = s1;
    red[threadIdx.x][1] = s2;
    __syncthreads();
    if (threadIdx.x < a.G * 2) {
        const int g = threadIdx.x >> 1, which = threadIdx.x & 1;
        const int lanes_per_group = 32 / a.G;      // channels of a lane never straddle a group
        double s = 0.0;
        for (int w = 0; w < kGWarps; ++w)
            for (int l = g * lanes_per_group; l < (g + 1) * lanes_per_group; ++l) s += red[w * 32 + l][which];
        a.part[(((size_t)b * gridDim.x + blockIdx.x) * a.G + g) * 2 + which] = s;
    }
}

// stats[b][g] = (mean, rstd); one warp per (b, g) sums the per-CTA fp64 partials in a fixed order
__global__ void gn_stats_kernel(const double *__restrict__ part, float *__restrict__ stats, int nblk, int G,
                                double count, float eps) {
    const int b = blockIdx.x, g = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (g >= G) return;
    double s1 = 0.0, s2 = 0.0;
    for (int i = lane; i < nblk; i += 32) {
        s1 += part[(((size_t)b * nblk + i) * G + g) * 2 + 0];
        s2 += part[(((size_t)b * nblk + i) * G + g) * 2 + 1];
    }
    for (int o = 16; o; o >>= 1) {
        s1 += __shfl_xor_sync(FULLM, s1, o);
        s2 += __shfl_xor_sync(FULLM, s2, o);
    }
    if (lane == 0) {
        double mean = s1 / count;
        double var = s2 / count - mean * mean;
        if (var < 0.0) var = 0.0;
        stats[((size_t)b * G + g) * 2 + 0] = (float)mean;
        stats[((size_t)b * G + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
    }
}

// GroupNorm + LeakyReLU of the selected extreme: out = LReLU((ysel - mean) * rstd * gamma + beta)
__global__ void edge_finish_kernel(const float *__restrict__ ysel, const float *__restrict__ stats,
                                   const float *__restrict__ gamma, const float *__restrict__ beta,
                                   float *__restrict__ out_nc, float *__restrict__ out_cn, int N, int Cout, int G,
                                   float slope) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int cpg = Cout / G;
    {
        const int c = c0 + threadIdx.x;
        const float gm = gamma[c], bt = beta[c];
        const float mean = stats[((size_t)b * G + c / cpg) * 2 + 0], rstd = stats[((size_t)b * G + c / cpg) * 2 + 1];
        for (int r = threadIdx.y; r < 32; r += blockDim.y) {
            const int n = n0 + r;
            float o = 0.f;
            if (n < N) {
                const size_t e = ((size_t)b * N + n) * Cout + c;
                o = lrelu((ysel[e] - mean) * rstd * gm + bt, slope);
                out_nc[e] = o;
            }
            tile[r][threadIdx.x] = o;
        }
    }
    if (out_cn == nullptr) return;
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int c = c0 + r, n = n0 + threadIdx.x;
        if (n < N) out_cn[((size_t)b * Cout + c) * N + n] = tile[threadIdx.x][r];
    }
}

// Same with 64 x 64 tiles and 16-byte accesses on all three streams (Cout % 64 == 0, N % 4 == 0).
__global__ void __launch_bounds__(256) edge_finish_wide_kernel(const float *__restrict__ ysel, const float *__restrict__ stats,
                                                               const float *__restrict__ gamma, const float *__restrict__ beta,
                                                               float *__restrict__ out_nc, float *__restrict__ out_cn, int N,
                                                               int Cout, int G, float slope) {
    __shared__ float tile[64][65];
    const int b = blockIdx.z;
    const int n0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
    const int cpg = Cout / G;
    {
        const int c = c0 + (threadIdx.x & 15) * 4;         // the same four channels in all four sweeps
        const float4 gm = __ldg(reinterpret_cast<const float4 *>(gamma + c)), bt = __ldg(reinterpret_cast<const float4 *>(beta + c));
        float mean[4], rstd[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            mean[j] = stats[((size_t)b * G + (c + j) / cpg) * 2 + 0];
            rstd[j] = stats[((size_t)b * G + (c + j) / cpg) * 2 + 1];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int n = (threadIdx.x >> 4) + 16 * i, cl = (threadIdx.x & 15) * 4;
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + n < N) {
                const size_t e = ((size_t)b * N + n0 + n) * Cout + c;
                const float4 y = __ldg(reinterpret_cast<const float4 *>(ysel + e));
                o.x = lrelu((y.x - mean[0]) * rstd[0] * gm.x + bt.x, slope);
                o.y = lrelu((y.y - mean[1]) * rstd[1] * gm.y + bt.y, slope);
                o.z = lrelu((y.z - mean[2]) * rstd[2] * gm.z + bt.z, slope);
                o.w = lrelu((y.w - mean[3]) * rstd[3] * gm.w + bt.w, slope);
                *reinterpret_cast<float4 *>(out_nc + e) = o;
            }
            tile[cl][n] = o.x; tile[cl + 1][n] = o.y; tile[cl + 2][n] = o.z; tile[cl + 3][n] = o.w;
        }
    }
    if (out_cn == nullptr) return;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int e = threadIdx.x + 256 * i, c = e >> 4, n = (e & 15) * 4;
        if (n0 + n < N)
            *reinterpret_cast<float4 *>(out_cn + ((size_t)b * Cout + c0 + c) * N + n0 + n) =
                make_float4(tile[c][n], tile[c][n + 1], tile[c][n + 2], tile[c][n + 3]);
    }
}

// ---------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------
// four consecutive elements of [P|Q], stored fp32 or bf16
__device__ __forceinline__ float4 ld4_pq(const void *base, size_t elem, int bf16) {
    if (bf16) {
        const uint2 w = __ldg(reinterpret_cast<const uint2 *>(reinterpret_cast<const __nv_bfloat16 *>(base) + elem));
        return make_float4(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xffff0000u), __uint_as_float(w.y << 16),
                           __uint_as_float(w.y & 0xffff0000u));
    }
    return *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(base) + elem);
}

struct BwdArgs {
    const void *pq;        // [B][N][2*Cout] fp32, or bf16 when pq_bf16
    const float *ysel, *ysum, *stats, *gamma, *beta, *gout;
    const unsigned char *arg;
    const int32_t *idx;
    float *part;       // [B][nblk][Cout][2]  per-CTA sums of du, du*yhat
    const float *coef; // [B][G][2] = (A_g, K_g)
    float *dpq;        // [B][N][2*Cout]
    int *deg;          // [B][N]
    int N, Cout, k, G;
    float slope;
    const float *x_nc; // [B][N][LDX]  small-C variant only
    float *xt;         // [B][N][LDX]  small-C variant only: sum of x_i over the in-edges (i -> j)
    int pq_bf16;
};

template <int VEC>
__global__ void __launch_bounds__(kGWarps * 32) edge_bwd_reduce_kernel(BwdArgs a) {
    __shared__ float red[kGWarps][32 * VEC][2];
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Cout = a.Cout, c0 = lane * VEC;
    const int cpg = Cout / a.G;
    const float mean = a.stats[((size_t)b * a.G + c0 / cpg) * 2 + 0], rstd = a.stats[((size_t)b * a.G + c0 / cpg) * 2 + 1];
    float gm[VEC], bt[VEC], s1[VEC], s2[VEC];
    VecIO<VEC>::ld(a.gamma + c0, gm);
    VecIO<VEC>::ld(a.beta + c0, bt);
#pragma unroll
    for (int v = 0; v < VEC; ++v) { s1[v] = 0.f; s2[v] = 0.f; }
    for (int pi = 0; pi < kPtsPerWarp; ++pi) {
        const int i = blockIdx.x * kPtsPerCta + warp * kPtsPerWarp + pi;
        if (i >= a.N) break;
        const size_t o = ((size_t)b * a.N + i) * Cout + c0;
        float ys[VEC], g[VEC];
        VecIO<VEC>::ld(a.ysel + o, ys);
        VecIO<VEC>::ld(a.gout + o, g);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            float yh = (ys[v] - mean) * rstd;
            float u = yh * gm[v] + bt[v];
            float du = u > 0.f ? g[v] : g[v] * a.slope;
            s1[v] += du;
            s2[v] = fmaf(du, yh, s2[v]);
        }
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) { red[warp][c0 + v][0] = s1[v]; red[warp][c0 + v][1] = s2[v]; }
    __syncthreads();
    for (int e = threadIdx.x; e < Cout * 2; e += blockDim.x) {
        const int c = e >> 1, which = e & 1;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kGWarps; ++w) s += red[w][c][which];
        a.part[(((size_t)b * gridDim.x + blockIdx.x) * Cout + c) * 2 + which] = s;
    }
}

// per cloud: S1[c] = sum_i du, S2[c] = sum_i du*yhat  ->  sbc[b][c][2] (double).  One warp per (cloud, channel)
// reduces the per-CTA partials in a fixed order (grid = (Cout / warps per CTA, B): fp64 adds are slow, spread them).
__global__ void edge_bwd_chansum_kernel(const float *__restrict__ part, double *__restrict__ sbc, int nblk, int Cout) {
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    const int c = blockIdx.x * nwarp + warp;
    if (c >= Cout) return;
    double s1 = 0.0, s2 = 0.0;
    for (int i = lane; i < nblk; i += 32) {
        const float2 v = *reinterpret_cast<const float2 *>(part + (((size_t)b * nblk + i) * Cout + c) * 2);
        s1 += (double)v.x;
        s2 += (double)v.y;
    }
    for (int o = 16; o; o >>= 1) {
        s1 += __shfl_xor_sync(FULLM, s1, o);
        s2 += __shfl_xor_sync(FULLM, s2, o);
    }
    if (lane == 0) {
        sbc[((size_t)b * Cout + c) * 2 + 0] = s1;
        sbc[((size_t)b * Cout + c) * 2 + 1] = s2;
    }
}

// coef[b][g] = (A_g, K_g) from the gamma-weighted channel sums of the group; one warp per (cloud, group)
__global__ void edge_bwd_coef_kernel(const double *__restrict__ sbc, const float *__restrict__ stats,
                                     const float *__restrict__ gamma, float *__restrict__ coef, int Cout, int G, double count) {
    const int b = blockIdx.x;
    const int g = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cpg = Cout / G;
    double m1 = 0.0, m2 = 0.0;
    for (int c = g * cpg + lane; c < (g + 1) * cpg; c += 32) {      // fixed order: lane-strided, then the shuffle tree
        m1 += sbc[((size_t)b * Cout + c) * 2 + 0] * (double)gamma[c];
        m2 += sbc[((size_t)b * Cout + c) * 2 + 1] * (double)gamma[c];
    }
    for (int o = 16; o; o >>= 1) {
        m1 += __shfl_xor_sync(FULLM, m1, o);
        m2 += __shfl_xor_sync(FULLM, m2, o);
    }
    if (lane == 0) {
        m1 /= count; m2 /= count;
        const double mean = stats[((size_t)b * G + g) * 2 + 0], rstd = stats[((size_t)b * G + g) * 2 + 1];
        coef[((size_t)b * G + g) * 2 + 0] = (float)(-rstd * m1 + rstd * rstd * m2 * mean);
        coef[((size_t)b * G + g) * 2 + 1] = (float)(-rstd * rstd * m2);
    }
}

// dgamma[c] = sum_b S2, dbeta[c] = sum_b S1
__global__ void edge_bwd_affine_kernel(const double *__restrict__ sbc, float *__restrict__ dgamma,
                                       float *__restrict__ dbeta, int B, int Cout) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= Cout) return;
    double s1 = 0.0, s2 = 0.0;
    for (int b = 0; b < B; ++b) { s1 += sbc[((size_t)b * Cout + c) * 2]; s2 += sbc[((size_t)b * Cout + c) * 2 + 1]; }
    dbeta[c] = (float)s1;
    dgamma[c] = (float)s2;
}

template <int VEC>
__global__ void __launch_bounds__(kGWarps * 32) edge_bwd_scatter_kernel(BwdArgs a) {
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Cout = a.Cout, k = a.k, c0 = lane * VEC;
    const int cpg = Cout / a.G;
    const int g = c0 / cpg;
    const float mean = a.stats[((size_t)b * a.G + g) * 2 + 0], rstd = a.stats[((size_t)b * a.G + g) * 2 + 1];
    const float Ag = a.coef[((size_t)b * a.G + g) * 2 + 0], Kg = a.coef[((size_t)b * a.G + g) * 2 + 1];
    float gm[VEC], bt[VEC];
    VecIO<VEC>::ld(a.gamma + c0, gm);
    VecIO<VEC>::ld(a.beta + c0, bt);
    float *dpq = a.dpq + (size_t)b * a.N * 2 * Cout;
    const size_t cloud0 = (size_t)b * a.N * 2 * Cout;
    // (the one-instruction byte addressing of the gather was tried here: the vector reductions then issue faster than the
    // L2 atomic units retire them and the kernel got SLOWER, 0.36 -> 0.51 ms; the typed form stays)
    float *dpq_c0 = dpq + c0;                              // 32-bit row offsets
    const unsigned row_stride = 2u * (unsigned)Cout;
    for (int pi = 0; pi < kPtsPerWarp; ++pi) {
        const int i = blockIdx.x * kPtsPerCta + warp * kPtsPerWarp + pi;
        if (i >= a.N) break;
        const size_t o = ((size_t)b * a.N + i) * Cout + c0;
        float ys[VEC], gg[VEC], ysum[VEC], q[VEC], s[VEC], T[VEC], dq[VEC];
        int ak[VEC];
        VecIO<VEC>::ld(a.ysel + o, ys);
        VecIO<VEC>::ld(a.gout + o, gg);
        VecIO<VEC>::ld(a.ysum + o, ysum);
        if (a.pq_bf16) ld_pq<VEC, true>(a.pq, cloud0 + (size_t)i * 2 * Cout + Cout + c0, q);
        else ld_pq<VEC, false>(a.pq, cloud0 + (size_t)i * 2 * Cout + Cout + c0, q);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            ak[v] = a.arg[o + v];
            float yh = (ys[v] - mean) * rstd;
            float u = yh * gm[v] + bt[v];
            float du = u > 0.f ? gg[v] : gg[v] * a.slope;
            s[v] = rstd * gm[v] * du;
            dq[v] = s[v] + (float)k * Ag + Kg * ysum[v];
            T[v] = fmaf(Kg, q[v], Ag);
        }
        VecIO<VEC>::st(dpq + (size_t)i * 2 * Cout + Cout + c0, dq);
        const int32_t *ip = a.idx + ((size_t)b * a.N + i) * k;
        for (int base = 0; base < k; base += 32) {
            const int cnt = min(32, k - base);
            const int myj = lane < cnt ? ip[base + lane] : 0;
            if (lane < cnt) atomicAdd(a.deg + (size_t)b * a.N + myj, 1);
            for (int t = 0; t < cnt; ++t) {
                const int j = __shfl_sync(FULLM, myj, t);
                float val[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) val[v] = T[v] + (ak[v] == base + t ? s[v] : 0.f);
                VecIO<VEC>::red(dpq_c0 + (unsigned)j * row_stride, val);
            }
        }
    }
}

// Small-C variant (LDX = 4 / 8: layer 1, xyz or xyz + normals) of the scatter pass.  The dense term of dP_j, the sum
// over in-edges of A_g + K_g Q_i, is linear in x_i (Q_i = Wq x_i): only X~_j = sum_{i->j} x_i (LDX floats per edge
// instead of Cout) and deg_j are scattered, and  deg_j A_g + K_g Wq X~_j  is added per point by
// edge_bwd_degfix_small_kernel.  The sparse term [k = k*] s_i is one scalar atomic per (point, channel).
// (The forward gather does not get the analogous treatment: rebuilding P_j = W1 x_j per edge costs more issue
// slots than the L2 traffic it saves -- measured 0.34 ms against 0.28 ms.)
template <int VEC, int LDX>
__global__ void __launch_bounds__(kGWarps * 32) edge_bwd_scatter_small_kernel(BwdArgs a) {
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Cout = a.Cout, k = a.k, c0 = lane * VEC;
    const int cpg = Cout / a.G;
    const int g = c0 / cpg;
    const float mean = a.stats[((size_t)b * a.G + g) * 2 + 0], rstd = a.stats[((size_t)b * a.G + g) * 2 + 1];
    const float Ag = a.coef[((size_t)b * a.G + g) * 2 + 0], Kg = a.coef[((size_t)b * a.G + g) * 2 + 1];
    float gm[VEC], bt[VEC];
    VecIO<VEC>::ld(a.gamma + c0, gm);
    VecIO<VEC>::ld(a.beta + c0, bt);
    float *dpq = a.dpq + (size_t)b * a.N * 2 * Cout;
    const float *xb = a.x_nc + (size_t)b * a.N * LDX;
    float *xt = a.xt + (size_t)b * a.N * LDX;
    for (int pi = 0; pi < kPtsPerWarp; ++pi) {
        const int i = blockIdx.x * kPtsPerCta + warp * kPtsPerWarp + pi;
        if (i >= a.N) break;
        const size_t o = ((size_t)b * a.N + i) * Cout + c0;
        float ys[VEC], gg[VEC], ysum[VEC], s[VEC], dq[VEC];
        VecIO<VEC>::ld(a.ysel + o, ys);
        VecIO<VEC>::ld(a.gout + o, gg);
        VecIO<VEC>::ld(a.ysum + o, ysum);
        const int32_t *ip = a.idx + ((size_t)b * a.N + i) * k;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const int ak = a.arg[o + v];
            float yh = (ys[v] - mean) * rstd;
            float u = yh * gm[v] + bt[v];
            float du = u > 0.f ? gg[v] : gg[v] * a.slope;
            s[v] = rstd * gm[v] * du;
            dq[v] = s[v] + (float)k * Ag + Kg * ysum[v];
            atomicAdd(dpq + (size_t)ip[ak] * 2 * Cout + c0 + v, s[v]);
        }
        VecIO<VEC>::st(dpq + (size_t)i * 2 * Cout + Cout + c0, dq);
        float xi[LDX];
#pragma unroll
        for (int c4 = 0; c4 < LDX / 4; ++c4) VecIO<4>::ld(xb + (size_t)i * LDX + c4 * 4, xi + c4 * 4);
        for (int base = 0; base < k; base += 32) {
            if (base + lane < k) {
                const int j = ip[base + lane];
                atomicAdd(a.deg + (size_t)b * a.N + j, 1);
#pragma unroll
                for (int c4 = 0; c4 < LDX / 4; ++c4) VecIO<4>::red(xt + (size_t)j * LDX + c4 * 4, xi + c4 * 4);
            }
        }
    }
}

// dP[j][c] += deg_j (A_g + K_g P[j][c]) + K_g sum_c' Wq[c'][c] X~_j[c'],  Wq[c'][c] = wcatT[Cout + c][c']
template <int LDX>
__global__ void edge_bwd_degfix_small_kernel(float *__restrict__ dpq, const void *__restrict__ pq, int pq_bf16, const int *__restrict__ deg,
                                             const float *__restrict__ coef, const float *__restrict__ xt,
                                             const float *__restrict__ wcatT, int N, int Cout, int G, long long total4) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total4) return;
    const int c4 = Cout / 4;
    long long bn = t / c4;
    int c = (int)(t % c4) * 4;
    int b = (int)(bn / N);
    const float Ag = coef[((size_t)b * G + c / (Cout / G)) * 2 + 0], Kg = coef[((size_t)b * G + c / (Cout / G)) * 2 + 1];
    const float dg = (float)deg[bn];
    float x[LDX];
#pragma unroll
    for (int q4 = 0; q4 < LDX / 4; ++q4) {
        float4 v = *reinterpret_cast<const float4 *>(xt + bn * LDX + q4 * 4);
        x[q4 * 4] = v.x; x[q4 * 4 + 1] = v.y; x[q4 * 4 + 2] = v.z; x[q4 * 4 + 3] = v.w;
    }
    float4 p = ld4_pq(pq, bn * 2 * Cout + c, pq_bf16);
    float4 *o = reinterpret_cast<float4 *>(dpq + bn * 2 * Cout + c);
    float4 v = *o;
    const float pv[4] = {p.x, p.y, p.z, p.w};
    float out[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float *wq = wcatT + (size_t)(Cout + c + e) * LDX;
        float qs = 0.f;
#pragma unroll
        for (int cc = 0; cc < LDX; ++cc) qs = fmaf(wq[cc], x[cc], qs);
        out[e] += dg * fmaf(Kg, pv[e], Ag) + Kg * qs;
    }
    *o = make_float4(out[0], out[1], out[2], out[3]);
}

// Small-C layer whose input needs NO gradient (layer 1: the input cloud is data): only dW is wanted, and
//   dWp[c] = sum_i sum_k dy_ikc x_j^T,   dy_ikc = [k = k*] s_ic + A_g + K_g (P_jc + Q_ic),   P_j = Wp x_j
// needs no scatter at all -- every term is a sum over the edges of a point or collapses onto per-cloud moments:
//   dWp[c] = sum_i s_ic x_{j*(i,c)}^T + K_g sum_i Q_ic Xbar_i^T + A_g (sum_i Xbar_i)^T + K_g (Wp S)[c],
//   Xbar_i = sum_k x_j,  S = sum_i sum_k x_j x_j^T (LDX x LDX per cloud),      dWq[c] = sum_i dQ_ic x_i^T.
// One warp per point stages its k neighbour rows (16 / 32 bytes each) in shared memory; nothing is written per point and
// there is no atomic: per-CTA partials of dWcat, summed in a fixed order by edge_bwd_dw_small_sum_kernel.  Replaces the
// X~ scatter, its degree fix-up, the [N][2 Cout] gradient buffer (and its memset) and the X^T dPQ product for this layer.
constexpr int kDwPts = 16;             // points per warp
template <int VEC, int LDX>
__global__ void __launch_bounds__(kGWarps * 32, LDX == 4 ? 3 : 1) edge_bwd_dw_small_kernel(BwdArgs a, const float *__restrict__ wcatT,
                                                                         float *__restrict__ part) {
    constexpr int COUT = 32 * VEC;
    __shared__ __align__(16) float s_red[kGWarps][2 * COUT][LDX];      // first the staged neighbour rows, then the reduction
    __shared__ float s_mom[kGWarps][LDX * LDX + LDX];
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = a.k, c0 = lane * VEC;
    const int cpg = COUT / a.G;
    const int g = c0 / cpg;
    const float mean = a.stats[((size_t)b * a.G + g) * 2 + 0], rstd = a.stats[((size_t)b * a.G + g) * 2 + 1];
    const float Ag = a.coef[((size_t)b * a.G + g) * 2 + 0], Kg = a.coef[((size_t)b * a.G + g) * 2 + 1];
    float gm[VEC], bt[VEC];
    VecIO<VEC>::ld(a.gamma + c0, gm);
    VecIO<VEC>::ld(a.beta + c0, bt);
    const float *xb = a.x_nc + (size_t)b * a.N * LDX;
    const size_t cloud0 = (size_t)b * a.N * 2 * COUT;
    float (*stage)[LDX] = reinterpret_cast<float (*)[LDX]>(&s_red[warp][0][0]);   // [64][LDX] of this warp (2 COUT >= 64 rows)
    float dwp[VEC][LDX], dwq[VEC][LDX], S[LDX][LDX], xtot[LDX];
#pragma unroll
    for (int cc = 0; cc < LDX; ++cc) {
        xtot[cc] = 0.f;
#pragma unroll
        for (int v = 0; v < VEC; ++v) { dwp[v][cc] = 0.f; dwq[v][cc] = 0.f; }
#pragma unroll
        for (int c2 = 0; c2 < LDX; ++c2) S[cc][c2] = 0.f;
    }
    // the neighbour ids of the NEXT point are fetched while the current one is processed: ids -> rows is a chain of two
    // dependent loads per point, and a warp walks its points one after the other
    const int i0 = (blockIdx.x * kGWarps + warp) * kDwPts;
    int jn[2] = {0, 0};
    if (i0 < a.N) {
#pragma unroll
        for (int t = 0; t < 2; ++t)
            if (lane + 32 * t < k) jn[t] = a.idx[((size_t)b * a.N + i0) * k + lane + 32 * t];
    }
    for (int pi = 0; pi < kDwPts; ++pi) {
        const int i = i0 + pi;
        if (i >= a.N) break;
        const int jc[2] = {jn[0], jn[1]};
        if (pi + 1 < kDwPts && i + 1 < a.N) {
#pragma unroll
            for (int t = 0; t < 2; ++t)
                if (lane + 32 * t < k) jn[t] = a.idx[((size_t)b * a.N + i + 1) * k + lane + 32 * t];
        }
        const size_t o = ((size_t)b * a.N + i) * COUT + c0;
        float ys[VEC], gg[VEC], ysum[VEC], q[VEC], sv[VEC], dq[VEC];
        int ak[VEC];
        VecIO<VEC>::ld(a.ysel + o, ys);
        VecIO<VEC>::ld(a.gout + o, gg);
        VecIO<VEC>::ld(a.ysum + o, ysum);
        if (a.pq_bf16) ld_pq<VEC, true>(a.pq, cloud0 + (size_t)i * 2 * COUT + COUT + c0, q);
        else ld_pq<VEC, false>(a.pq, cloud0 + (size_t)i * 2 * COUT + COUT + c0, q);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            ak[v] = a.arg[o + v];
            const float yh = (ys[v] - mean) * rstd;
            const float u = yh * gm[v] + bt[v];
            const float du = u > 0.f ? gg[v] : gg[v] * a.slope;
            sv[v] = rstd * gm[v] * du;
            dq[v] = sv[v] + (float)k * Ag + Kg * ysum[v];
        }
        float xi[LDX], xs[LDX];
#pragma unroll
        for (int c4 = 0; c4 < LDX / 4; ++c4) VecIO<4>::ld(xb + (size_t)i * LDX + c4 * 4, xi + c4 * 4);
#pragma unroll
        for (int cc = 0; cc < LDX; ++cc) xs[cc] = 0.f;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int kk = lane + 32 * t;
            if (kk < k) {
                const int j = jc[t];
                float xj[LDX];
#pragma unroll
                for (int c4 = 0; c4 < LDX / 4; ++c4) {
                    VecIO<4>::ld(xb + (size_t)j * LDX + c4 * 4, xj + c4 * 4);
                    VecIO<4>::st(&stage[kk][c4 * 4], xj + c4 * 4);
                }
#pragma unroll
                for (int cc = 0; cc < LDX; ++cc) {
                    xs[cc] += xj[cc];
#pragma unroll
                    for (int c2 = 0; c2 < LDX; ++c2) S[cc][c2] = fmaf(xj[cc], xj[c2], S[cc][c2]);
                }
            }
        }
#pragma unroll
        for (int cc = 0; cc < LDX; ++cc) {
            for (int off = 16; off; off >>= 1) xs[cc] += __shfl_xor_sync(FULLM, xs[cc], off);
            xtot[cc] += xs[cc];
        }
        __syncwarp();
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const float kq = Kg * q[v];
#pragma unroll
            for (int cc = 0; cc < LDX; ++cc) {
                dwp[v][cc] = fmaf(sv[v], stage[ak[v]][cc], fmaf(kq, xs[cc], dwp[v][cc]));
                dwq[v][cc] = fmaf(dq[v], xi[cc], dwq[v][cc]);
            }
        }
        __syncwarp();
    }
    // per-warp moments: S over the lanes (every lane saw different edges), sum of Xbar (already uniform)
#pragma unroll
    for (int cc = 0; cc < LDX; ++cc)
#pragma unroll
        for (int c2 = 0; c2 < LDX; ++c2) {
            float v = S[cc][c2];
            for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(FULLM, v, off);
            if (lane == 0) s_mom[warp][cc * LDX + c2] = v;
        }
    if (lane == 0) {
#pragma unroll
        for (int cc = 0; cc < LDX; ++cc) s_mom[warp][LDX * LDX + cc] = xtot[cc];
    }
    __syncthreads();                                   // every warp is done with its staging rows
#pragma unroll
    for (int v = 0; v < VEC; ++v)
#pragma unroll
        for (int cc = 0; cc < LDX; ++cc) { s_red[warp][c0 + v][cc] = dwp[v][cc]; s_red[warp][COUT + c0 + v][cc] = dwq[v][cc]; }
    __syncthreads();
    float *out = part + ((size_t)b * gridDim.x + blockIdx.x) * 2 * COUT * LDX;
    for (int e = threadIdx.x; e < 2 * COUT * LDX; e += blockDim.x) {
        const int n = e / LDX, cc = e % LDX;
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < kGWarps; ++w) v += s_red[w][n][cc];
        if (n < COUT) {                                // cloud-level terms of dWp, from this CTA's share of the moments
            const int gg2 = n / cpg;
            const float A2 = a.coef[((size_t)b * a.G + gg2) * 2 + 0], K2 = a.coef[((size_t)b * a.G + gg2) * 2 + 1];
            float xt = 0.f, ws = 0.f;
            for (int w = 0; w < kGWarps; ++w) {
                xt += s_mom[w][LDX * LDX + cc];
#pragma unroll
                for (int c2 = 0; c2 < LDX; ++c2) ws = fmaf(wcatT[(size_t)n * LDX + c2], s_mom[w][c2 * LDX + cc], ws);
            }
            v += A2 * xt + K2 * ws;
        }
        out[e] = v;
    }
}

// dwcat[cc][n] = sum over the CTAs of part[cta][n][cc], fixed order.  block (32 entries, 32 slices of the CTA list)
__global__ void __launch_bounds__(1024) edge_bwd_dw_small_sum_kernel(const float *__restrict__ part, float *__restrict__ dwcat,
                                                                     int ctas, int n2, int ldx) {
    __shared__ double red[32][33];
    const int e = blockIdx.x * 32 + threadIdx.x;
    double s = 0.0;
    if (e < n2 * ldx)
        for (int i = threadIdx.y; i < ctas; i += 32) s += (double)part[(size_t)i * n2 * ldx + e];
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && e < n2 * ldx) {
        double t = 0.0;
        for (int r = 0; r < 32; ++r) t += red[r][threadIdx.x];
        dwcat[(size_t)(e % ldx) * n2 + e / ldx] = (float)t;
    }
}

// C = 64 feeding Cout > 64 (layer 3): the same split as the small-C variant, with the X~ scatter as one 16-byte
// reduction per lane -- each half-warp owns one edge (16 lanes x 4 floats = the 64-float row of x_i) -- so an edge costs
// 256 B of L2 reductions instead of 4 Cout, and the dense term comes back as one tensor-core GEMM  X~ Wq^T  whose
// result edge_bwd_degfix_mid_kernel folds in.
template <int VEC>
__global__ void __launch_bounds__(kGWarps * 32) edge_bwd_scatter_mid_kernel(BwdArgs a) {
    constexpr int LDX = 64;
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Cout = a.Cout, k = a.k, c0 = lane * VEC;
    const int cpg = Cout / a.G;
    const int g = c0 / cpg;
    const float mean = a.stats[((size_t)b * a.G + g) * 2 + 0], rstd = a.stats[((size_t)b * a.G + g) * 2 + 1];
    const float Ag = a.coef[((size_t)b * a.G + g) * 2 + 0], Kg = a.coef[((size_t)b * a.G + g) * 2 + 1];
    float gm[VEC], bt[VEC];
    VecIO<VEC>::ld(a.gamma + c0, gm);
    VecIO<VEC>::ld(a.beta + c0, bt);
    float *dpq = a.dpq + (size_t)b * a.N * 2 * Cout;
    const float *xb = a.x_nc + (size_t)b * a.N * LDX;
    const int half = lane >> 4;
    float *xt_l = a.xt + (size_t)b * a.N * LDX + (lane & 15) * 4;
    float *dpq_c0 = dpq + c0;
    const unsigned row_stride = 2u * (unsigned)Cout;
    for (int pi = 0; pi < kPtsPerWarp; ++pi) {
        const int i = blockIdx.x * kPtsPerCta + warp * kPtsPerWarp + pi;
        if (i >= a.N) break;
        const size_t o = ((size_t)b * a.N + i) * Cout + c0;
        float ys[VEC], gg[VEC], ysum[VEC], s[VEC], dq[VEC];
        VecIO<VEC>::ld(a.ysel + o, ys);
        VecIO<VEC>::ld(a.gout + o, gg);
        VecIO<VEC>::ld(a.ysum + o, ysum);
        const int32_t *ip = a.idx + ((size_t)b * a.N + i) * k;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const int ak = a.arg[o + v];
            float yh = (ys[v] - mean) * rstd;
            float u = yh * gm[v] + bt[v];
            float du = u > 0.f ? gg[v] : gg[v] * a.slope;
            s[v] = rstd * gm[v] * du;
            dq[v] = s[v] + (float)k * Ag + Kg * ysum[v];
            atomicAdd(dpq_c0 + (unsigned)ip[ak] * row_stride + v, s[v]);
        }
        VecIO<VEC>::st(dpq + (size_t)i * 2 * Cout + Cout + c0, dq);
        float xi[4];
        VecIO<4>::ld(xb + (size_t)i * LDX + (lane & 15) * 4, xi);
        for (int base = 0; base < k; base += 32) {
            const int cnt = min(32, k - base);
            const int myj = lane < cnt ? ip[base + lane] : 0;
            if (lane < cnt) atomicAdd(a.deg + (size_t)b * a.N + myj, 1);
            for (int t = 0; t < cnt; t += 2) {
                const int j = __shfl_sync(FULLM, myj, t + half);
                if (t + half < cnt) VecIO<4>::red(xt_l + (unsigned)j * LDX, xi);
            }
        }
    }
}

// dP[j][c] += deg_j (A_g + K_g P[j][c]) + K_g (X~ Wq^T)[j][c]
__global__ void edge_bwd_degfix_mid_kernel(float *__restrict__ dpq, const void *__restrict__ pq, int pq_bf16, const int *__restrict__ deg,
                                           const float *__restrict__ coef, const float *__restrict__ xq, int N, int Cout,
                                           int G, long long total4) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total4) return;
    const int c4 = Cout / 4;
    long long bn = t / c4;
    int c = (int)(t % c4) * 4;
    int b = (int)(bn / N);
    const float Ag = coef[((size_t)b * G + c / (Cout / G)) * 2 + 0], Kg = coef[((size_t)b * G + c / (Cout / G)) * 2 + 1];
    const float dg = (float)deg[bn];
    float4 p = ld4_pq(pq, bn * 2 * Cout + c, pq_bf16);
    float4 x = *reinterpret_cast<const float4 *>(xq + bn * Cout + c);
    float4 *o = reinterpret_cast<float4 *>(dpq + bn * 2 * Cout + c);
    float4 v = *o;
    v.x += dg * fmaf(Kg, p.x, Ag) + Kg * x.x;
    v.y += dg * fmaf(Kg, p.y, Ag) + Kg * x.y;
    v.z += dg * fmaf(Kg, p.z, Ag) + Kg * x.z;
    v.w += dg * fmaf(Kg, p.w, Ag) + Kg * x.w;
    *o = v;
}

// dP[j][c] += deg_j * K_g * P[j][c]
__global__ void edge_bwd_degfix_kernel(float *__restrict__ dpq, const void *__restrict__ pq, int pq_bf16, const int *__restrict__ deg,
                                       const float *__restrict__ coef, int N, int Cout, int G, long long total4) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total4) return;
    const int c4 = Cout / 4;
    long long bn = t / c4;
    int c = (int)(t % c4) * 4;
    int b = (int)(bn / N);
    const float Kg = coef[((size_t)b * G + c / (Cout / G)) * 2 + 1];
    const float d = (float)deg[bn] * Kg;
    float4 p = ld4_pq(pq, bn * 2 * Cout + c, pq_bf16);
    float4 *o = reinterpret_cast<float4 *>(dpq + bn * 2 * Cout + c);
    float4 v = *o;
    v.x = fmaf(d, p.x, v.x); v.y = fmaf(d, p.y, v.y); v.z = fmaf(d, p.z, v.z); v.w = fmaf(d, p.w, v.w);
    *o = v;
}

// ---------------------------------------------------------------------------------
// buffer plans
// ---------------------------------------------------------------------------------
struct Saved {
    float *pq, *ysel, *ysum, *stats;      // pq: fp32, or bf16 (half the bytes) when d->storage_bf16
    unsigned char *arg;
};

static size_t plan_saved(const gcanet_edgeconv_desc *d, void *base, Saved *s) {
    Carver cv(base);
    size_t bn = (size_t)d->B * d->N;
    float *pq = d->storage_bf16 ? reinterpret_cast<float *>(cv.take<__nv_bfloat16>(bn * 2 * d->Cout)) : cv.take<float>(bn * 2 * d->Cout);
    float *ysel = cv.take<float>(bn * d->Cout);
    float *ysum = cv.take<float>(bn * d->Cout);
    unsigned char *arg = cv.take<unsigned char>(bn * d->Cout);
    float *stats = cv.take<float>((size_t)d->B * d->groups * 2);
    if (s) { s->pq = pq; s->ysel = ysel; s->ysum = ysum; s->arg = arg; s->stats = stats; }
    return cv.off;
}

struct FwdWs {
    float *wcat, *wcatT;
    double *part;
    float *pq32;       // bf16 storage: fp32 staging of [P|Q] where the projection runs on the CUDA cores (xyz layer)
};

static size_t plan_fwd(const gcanet_edgeconv_desc *d, void *base, FwdWs *w) {
    Carver cv(base);
    int nblk = ceil_div(d->N, kPtsPerCta);
    float *wcat = cv.take<float>((size_t)d->ldx * 2 * d->Cout);
    float *wcatT = cv.take<float>((size_t)d->ldx * 2 * d->Cout);
    double *part = cv.take<double>((size_t)d->B * nblk * d->groups * 2);
    float *pq32 = cv.take<float>(d->storage_bf16 ? (size_t)d->B * d->N * 2 * d->Cout : 0);
    if (w) { w->wcat = wcat; w->wcatT = wcatT; w->part = part; w->pq32 = pq32; }
    return cv.off;
}

struct BwdWs {
    float *wcatT, *wcat, *dpq, *part, *coef, *dwcat, *dwpart, *xt, *xq, *dwsmall;
    int *deg;
    double *sbc;
};

// layer-3 shape: 64 input channels feeding more output channels -- scatter X~ instead of the Cout-wide dense term
// (at Cout = 64 the split loses: 0.65 ms against 0.55 ms for the layer's backward)
static bool bwd_mid_path(const gcanet_edgeconv_desc *d) {
    static const bool off = GCANET_AID_ENV("GCANET_NO_MID_SCATTER") != nullptr;
    return !off && d->ldx == 64 && d->C == 64 && d->Cout >= 128 && d->Cout <= 256 && (long long)d->B * d->N >= 1024;
}

static size_t plan_bwd(const gcanet_edgeconv_desc *d, void *base, BwdWs *w) {
    Carver cv(base);
    size_t bn = (size_t)d->B * d->N;
    int nblk = ceil_div(d->N, kPtsPerCta);
    int M = (int)bn;
    float *wcatT = cv.take<float>((size_t)d->ldx * 2 * d->Cout);
    float *dpq = cv.take<float>(bn * 2 * d->Cout);
    int *deg = cv.take<int>(bn);
    float *part = cv.take<float>((size_t)d->B * nblk * d->Cout * 2);
    double *sbc = cv.take<double>((size_t)d->B * d->Cout * 2);
    float *coef = cv.take<float>((size_t)d->B * d->groups * 2);
    float *dwcat = cv.take<float>((size_t)d->ldx * 2 * d->Cout);
    float *dwpart = cv.take<float>((size_t)tn_splits(M, 2 * d->Cout, d->ldx) * d->ldx * 2 * d->Cout);
    const bool mid = bwd_mid_path(d);
    float *xt = cv.take<float>(d->ldx <= 8 || mid ? bn * d->ldx : 0);
    float *xq = cv.take<float>(mid ? bn * d->Cout : 0);
    float *wcat = cv.take<float>((size_t)d->ldx * 2 * d->Cout);
    // per-CTA partials of the scatter-free weight gradient (small-C layers whose input needs no gradient)
    float *dwsmall = cv.take<float>(d->ldx <= 8 && d->Cout <= 64 ? (size_t)d->B * ceil_div(d->N, kGWarps * kDwPts) * 2 * d->Cout * d->ldx : 0);
    if (w) { w->dwsmall = dwsmall; w->xt = xt; w->xq = xq; w->wcat = wcat; w->wcatT = wcatT; w->dpq = dpq; w->deg = deg; w->part = part; w->sbc = sbc; w->coef = coef; w->dwcat = dwcat; w->dwpart = dwpart; }
    return cv.off;
}

static int check_desc(const gcanet_edgeconv_desc *d) {
    GCANET_REQUIRE(d != nullptr, "edgeconv: null descriptor");
    GCANET_REQUIRE(d->B >= 1 && d->N >= 1 && d->C >= 1 && d->k >= 1, "edgeconv: bad shape B=%d N=%d C=%d k=%d", d->B, d->N, d->C, d->k);
    GCANET_REQUIRE(d->B <= 65535, "edgeconv: B > 65535");
    GCANET_REQUIRE(d->k <= 255, "edgeconv: k=%d > 255 (arg index is one byte)", d->k);
    GCANET_REQUIRE(d->ldx >= d->C && d->ldx % 4 == 0 && d->ldx <= 256, "edgeconv: ldx=%d must be a multiple of 4 in [C, 256]", d->ldx);
    GCANET_REQUIRE(d->Cout % 32 == 0 && d->Cout >= 32 && d->Cout <= 256, "edgeconv: Cout=%d must be a multiple of 32 in [32, 256]", d->Cout);
    int vec = d->Cout / 32;
    GCANET_REQUIRE(vec == 1 || vec == 2 || vec == 4 || vec == 8, "edgeconv: Cout=%d unsupported (Cout/32 must be 1, 2, 4 or 8)", d->Cout);
    GCANET_REQUIRE(d->groups >= 1 && d->Cout % d->groups == 0 && 32 % d->groups == 0,
                   "edgeconv: groups=%d must divide 32 and Cout=%d", d->groups, d->Cout);
    GCANET_REQUIRE(d->eps > 0.f, "edgeconv: eps must be positive");
    // max_k LReLU(GN(y)) = LReLU(GN(max or min of y)) needs a non-decreasing activation
    GCANET_REQUIRE(d->slope >= 0.f, "edgeconv: negative_slope=%g < 0 (the fused max/min selection needs a non-decreasing activation)", (double)d->slope);
    GCANET_REQUIRE((long long)d->B * d->N < 2147483647ll, "edgeconv: B * N = %lld does not fit 32-bit row counts", (long long)d->B * d->N);
    GCANET_REQUIRE((long long)d->N * 2 * d->Cout < (1ll << 30), "edgeconv: N * 2 * Cout = %lld exceeds 2^30 (32-bit row offsets)",
                   (long long)d->N * 2 * d->Cout);
    return GCANET_OK;
}

// [P|Q] staging for the bf16-storage mode (xyz layer, whose projection runs on the CUDA cores): four floats -> four bf16
__global__ void f32_to_bf16_kernel(const float *__restrict__ src, __nv_bfloat16 *__restrict__ dst, long long n4) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 v = __ldg(reinterpret_cast<const float4 *>(src) + i);
    const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    reinterpret_cast<uint2 *>(dst)[i] = make_uint2(*reinterpret_cast<const uint32_t *>(&a), *reinterpret_cast<const uint32_t *>(&b));
}
template <int VEC>
static int run_forward(const gcanet_edgeconv_desc *d, const float *x_nc, const int32_t *idx, const float *weight,
                       const float *gamma, const float *beta, float *out_nc, float *out_cn, const Saved &sv,
                       const FwdWs &w, cudaStream_t st) {
    const int M = d->B * d->N, Cout = d->Cout;
    int total = d->ldx * 2 * Cout;
    prep_wcat_kernel<<<ceil_div(total, 256), 256, 0, st>>>(weight, w.wcat, w.wcatT, d->C, d->ldx, Cout);
    GCANET_LAUNCH_OK("prep_wcat_kernel");
    // [P|Q] = X Wcat: tensor cores (bf16x3 split, fp32-accurate) for the feature layers, CUDA cores for the xyz layer
    int rc = gemm_tc_try(x_nc, d->ldx, w.wcatT, d->ldx, sv.pq, 2 * Cout, M, 2 * Cout, d->ldx, st, d->storage_bf16);
    if (rc > 0) {
        float *dst = d->storage_bf16 ? w.pq32 : sv.pq;
        rc = launch_sgemm_nn(x_nc, w.wcat, dst, M, 2 * Cout, d->ldx, d->ldx, 2 * Cout, 2 * Cout, st);
        if (rc == GCANET_OK && d->storage_bf16) {
            const long long n4 = (long long)M * 2 * Cout / 4;
            f32_to_bf16_kernel<<<(unsigned)ceil_div64(n4, 256), 256, 0, st>>>(w.pq32, reinterpret_cast<__nv_bfloat16 *>(sv.pq), n4);
            GCANET_LAUNCH_OK("f32_to_bf16_kernel");
        }
    }
    if (rc) return rc;
    FwdArgs fa{sv.pq, idx, gamma, sv.ysel, sv.ysum, sv.arg, w.part, d->N, Cout, d->k, d->groups};
    const int nblk = ceil_div(d->N, kPtsPerCta);
    if (d->storage_bf16) edge_gather_reduce_kernel<VEC, true><<<dim3(nblk, d->B), kGWarps * 32, 0, st>>>(fa);
    else edge_gather_reduce_kernel<VEC, false><<<dim3(nblk, d->B), kGWarps * 32, 0, st>>>(fa);
    GCANET_LAUNCH_OK("edge_gather_reduce_kernel");
    double count = (double)(Cout / d->groups) * d->N * d->k;
    gn_stats_kernel<<<d->B, 32 * d->groups, 0, st>>>(w.part, sv.stats, nblk, d->groups, count, d->eps);
    GCANET_LAUNCH_OK("gn_stats_kernel");
    const uintptr_t align_bits = reinterpret_cast<uintptr_t>(out_nc) | reinterpret_cast<uintptr_t>(out_cn) |
                                 reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta);
    if (Cout % 64 == 0 && d->N % 4 == 0 && (align_bits & 15) == 0) {
        edge_finish_wide_kernel<<<dim3(ceil_div(d->N, 64), Cout / 64, d->B), 256, 0, st>>>(sv.ysel, sv.stats, gamma, beta, out_nc,
                                                                                         out_cn, d->N, Cout, d->groups, d->slope);
        GCANET_LAUNCH_OK("edge_finish_wide_kernel");
    } else {
        dim3 fg(ceil_div(d->N, 32), Cout / 32, d->B), fb(32, 8);
        edge_finish_kernel<<<fg, fb, 0, st>>>(sv.ysel, sv.stats, gamma, beta, out_nc, out_cn, d->N, Cout, d->groups, d->slope);
        GCANET_LAUNCH_OK("edge_finish_kernel");
    }
    return GCANET_OK;
}

template <int VEC>
static int run_backward(const gcanet_edgeconv_desc *d, const float *x_nc, const int32_t *idx, const float *weight,
                        const float *gamma, const float *beta, const float *gout, const Saved &sv, float *grad_x_nc,
                        float *grad_weight, float *grad_gamma, float *grad_beta, const BwdWs &w, cudaStream_t st) {
    const int M = d->B * d->N, Cout = d->Cout;
    const size_t bn = (size_t)M;
    const int nblk = ceil_div(d->N, kPtsPerCta);
    int total = d->ldx * 2 * Cout;
    prep_wcat_kernel<<<ceil_div(total, 256), 256, 0, st>>>(weight, w.wcat, w.wcatT, d->C, d->ldx, Cout);
    GCANET_LAUNCH_OK("prep_wcat_kernel");
    const bool small = d->ldx == 4 || d->ldx == 8;
    const bool mid = bwd_mid_path(d);
    // layer 1 in training: the input cloud is data, so only dW is needed and it can be had without any scatter
    const bool dw_only = small && grad_x_nc == nullptr && VEC <= 2 && d->k <= 64 && !GCANET_AID_ENV("GCANET_NO_DW_ONLY");
    if (!dw_only) {
        GCANET_CUDA_OK(cudaMemsetAsync(w.dpq, 0, bn * 2 * Cout * sizeof(float), st));
        GCANET_CUDA_OK(cudaMemsetAsync(w.deg, 0, bn * sizeof(int), st));
        if (small || mid) GCANET_CUDA_OK(cudaMemsetAsync(w.xt, 0, bn * d->ldx * sizeof(float), st));
    }
    const void *pq32 = sv.pq;                 // the backward kernels read [P|Q] per point, in the precision it was saved in
    const int pqb = d->storage_bf16;
    BwdArgs ba{pq32, sv.ysel, sv.ysum, sv.stats, gamma, beta, gout, sv.arg, idx, w.part, w.coef, w.dpq, w.deg,
               d->N, Cout, d->k, d->groups, d->slope, x_nc, w.xt, pqb};
    edge_bwd_reduce_kernel<VEC><<<dim3(nblk, d->B), kGWarps * 32, 0, st>>>(ba);
    GCANET_LAUNCH_OK("edge_bwd_reduce_kernel");
    double count = (double)(Cout / d->groups) * d->N * d->k;
    edge_bwd_chansum_kernel<<<dim3(ceil_div(Cout, 8), d->B), 256, 0, st>>>(w.part, w.sbc, nblk, Cout);
    GCANET_LAUNCH_OK("edge_bwd_chansum_kernel");
    edge_bwd_coef_kernel<<<d->B, 32 * d->groups, 0, st>>>(w.sbc, sv.stats, gamma, w.coef, Cout, d->groups, count);
    GCANET_LAUNCH_OK("edge_bwd_coef_kernel");
    edge_bwd_affine_kernel<<<ceil_div(Cout, 128), 128, 0, st>>>(w.sbc, grad_gamma, grad_beta, d->B, Cout);
    GCANET_LAUNCH_OK("edge_bwd_affine_kernel");
    long long total4 = (long long)bn * (Cout / 4);
    if (dw_only) {
        if constexpr (VEC <= 2) {
            const int ctas = ceil_div(d->N, kGWarps * kDwPts);
            if (d->ldx == 4) edge_bwd_dw_small_kernel<VEC, 4><<<dim3(ctas, d->B), kGWarps * 32, 0, st>>>(ba, w.wcatT, w.dwsmall);
            else edge_bwd_dw_small_kernel<VEC, 8><<<dim3(ctas, d->B), kGWarps * 32, 0, st>>>(ba, w.wcatT, w.dwsmall);
            GCANET_LAUNCH_OK("edge_bwd_dw_small_kernel");
            edge_bwd_dw_small_sum_kernel<<<ceil_div(2 * Cout * d->ldx, 32), dim3(32, 32), 0, st>>>(w.dwsmall, w.dwcat, ctas * d->B,
                                                                                                 2 * Cout, d->ldx);
            GCANET_LAUNCH_OK("edge_bwd_dw_small_sum_kernel");
            unprep_dw_kernel<<<ceil_div(Cout * d->C, 256), 256, 0, st>>>(w.dwcat, grad_weight, d->C, Cout);
            GCANET_LAUNCH_OK("unprep_dw_kernel");
        }
        return GCANET_OK;
    }
    if (small) {
        if (d->ldx == 4) edge_bwd_scatter_small_kernel<VEC, 4><<<dim3(nblk, d->B), kGWarps * 32, 0, st>>>(ba);
        else edge_bwd_scatter_small_kernel<VEC, 8><<<dim3(nblk, d->B), kGWarps * 32, 0, st>>>(ba);
        GCANET_LAUNCH_OK("edge_bwd_scatter_small_kernel");
        if (d->ldx == 4)
            edge_bwd_degfix_small_kernel<4><<<(unsigned)ceil_div64(total4, 256), 256, 0, st>>>(w.dpq, pq32, pqb, w.deg, w.coef, w.xt, w.wcatT,
                                                                                           d->N, Cout, d->groups, total4);
        else
            edge_bwd_degfix_small_kernel<8><<<(unsigned)ceil_div64(total4, 256), 256, 0, st>>>(w.dpq, pq32, pqb, w.deg, w.coef, w.xt, w.wcatT,
                                                                                           d->N, Cout, d->groups, total4);
        GCANET_LAUNCH_OK("edge_bwd_degfix_small_kernel");
    } else if (mid) {
        if constexpr (VEC >= 4) {
            edge_bwd_scatter_mid_kernel<VEC><<<dim3(nblk, d->B), kGWarps * 32, 0, st>>>(ba);
            GCANET_LAUNCH_OK("edge_bwd_scatter_mid_kernel");
            // X~ Wq^T: rows Cout.. of wcatT are Wq
            int rq = gemm_tc_try(w.xt, d->ldx, w.wcatT + (size_t)Cout * d->ldx, d->ldx, w.xq, Cout, M, Cout, d->ldx, st);
            if (rq > 0) rq = launch_sgemm_nn(w.xt, w.wcat + Cout, w.xq, M, Cout, d->ldx, d->ldx, 2 * Cout, Cout, st);
            if (rq) return rq;
            edge_bwd_degfix_mid_kernel<<<(unsigned)ceil_div64(total4, 256), 256, 0, st>>>(w.dpq, pq32, pqb, w.deg, w.coef, w.xq, d->N,
                                                                                        Cout, d->groups, total4);
            GCANET_LAUNCH_OK("edge_bwd_degfix_mid_kernel");
        }
    } else {
        edge_bwd_scatter_kernel<VEC><<<dim3(nblk, d->B), kGWarps * 32, 0, st>>>(ba);
        GCANET_LAUNCH_OK("edge_bwd_scatter_kernel");
        edge_bwd_degfix_kernel<<<(unsigned)ceil_div64(total4, 256), 256, 0, st>>>(w.dpq, pq32, pqb, w.deg, w.coef, d->N, Cout,
                                                                                d->groups, total4);
        GCANET_LAUNCH_OK("edge_bwd_degfix_kernel");
    }
    // dWcat[ldx][2Cout] = X^T dPQ ; dW from it
    int rc = launch_sgemm_tn(x_nc, w.dpq, w.dwcat, w.dwpart, M, 2 * Cout, d->ldx, d->ldx, 2 * Cout, st);
    if (rc) return rc;
    unprep_dw_kernel<<<ceil_div(Cout * d->C, 256), 256, 0, st>>>(w.dwcat, grad_weight, d->C, Cout);
    GCANET_LAUNCH_OK("unprep_dw_kernel");
    if (grad_x_nc) {
        rc = gemm_tc_try(w.dpq, 2 * Cout, w.wcat, 2 * Cout, grad_x_nc, d->ldx, M, d->ldx, 2 * Cout, st);
        if (rc > 0) rc = launch_sgemm_nn(w.dpq, w.wcatT, grad_x_nc, M, d->ldx, 2 * Cout, 2 * Cout, d->ldx, d->ldx, st);
        if (rc) return rc;
    }
    return GCANET_OK;
}

// =================================================================================
// EdgeConv on normals (conv_normal, M4:584-587 / M4:691-693): 7-channel edge feature
//   e_ik = (clamp(n_i.n_j, -.99, .99), n_j - n_i, n_i),  y_ik = W e_ik,  then GroupNorm, LeakyReLU, max over k.
// The feature is rebuilt per edge from the 12-byte normal of the neighbour, so neither
// [B][7][N][k] nor [B][64][N][k] exists.  Only parameter gradients are produced: the inputs of this
// head are data (points and normals), never activations.  With dy_ik = [k=k*] s + A_g + K_g y_ik:
//   dW[c][f] = sum_i s_ic e_{i,k*(i,c),f}  +  sum_b ( A_g E1_b[f] + K_g sum_f' W[c][f'] E2_b[f'][f] ),
// E1_b = sum_ik e_ik, E2_b = sum_ik e_ik e_ik^T (35 numbers per cloud, accumulated in the forward pass).
// =================================================================================
constexpr int kNF = 7;                 // edge-feature channels
constexpr int kNMom = 35;              // 7 first moments + 28 unique second moments

__device__ __forceinline__ float normal_angle(float a0, float a1, float a2, float b0, float b1, float b2) {
    // reference: elementwise product summed over the 3 channels, then clamp (M4:194)
    float d = __fadd_rn(__fadd_rn(__fmul_rn(a0, b0), __fmul_rn(a1, b1)), __fmul_rn(a2, b2));
    return fminf(fmaxf(d, -0.99f), 0.99f);
}

struct NFwdArgs {
    const float *x_nc;     // [B][N][ldx], normals in columns 3..5
    const int32_t *idx;    // [B][N][k]
    const float *weight;   // [Cout][7]
    const float *gamma;
    float *ysel;           // [B][N][Cout]
    unsigned char *arg;    // [B][N][Cout]
    double *part;          // [B][nblk][G][2]   group moments of y
    double *mom_part;      // [B][nblk][kNMom]  feature moments
    int N, ldx, Cout, k, G;
};

template <int VEC>
__global__ void __launch_bounds__(kGWarps * 32) normal_edge_forward_kernel(NFwdArgs a) {
    __shared__ double red[kGWarps * 32][2];
    __shared__ float4 s_edge[kGWarps][64];
    __shared__ double s_mom[kGWarps][kNMom];
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Cout = a.Cout, k = a.k, c0 = lane * VEC;
    const float *xb = a.x_nc + (size_t)b * a.N * a.ldx;
    float w[VEC][kNF], sg[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
#pragma unroll
        for (int f = 0; f < kNF; ++f) w[v][f] = a.weight[(size_t)(c0 + v) * kNF + f];
        sg[v] = a.gamma[c0 + v] < 0.f ? -1.f : 1.f;
    }
    // Edges are staged 64 at a time: lane l loads the neighbour index and normal of edges l and 32 + l (two independent
    // loads per lane instead of a chain of k dependent ones), forms (angle, n_j) and adds its edges' 35 feature moments
    // (7 first + 28 second, items in row-major order of the pairs f <= g) to its own registers; the channel loop then
    // reads the staged edges back as broadcasts.  The moments are summed over the lanes once, at the end.
    float macc[kNMom];
#pragma unroll
    for (int m = 0; m < kNMom; ++m) macc[m] = 0.f;
    double s1 = 0.0, s2 = 0.0;

    for (int pi = 0; pi < kPtsPerWarp; ++pi) {
        const int i = blockIdx.x * kPtsPerCta + warp * kPtsPerWarp + pi;
        if (i >= a.N) break;
        const float a0 = xb[(size_t)i * a.ldx + 3], a1 = xb[(size_t)i * a.ldx + 4], a2 = xb[(size_t)i * a.ldx + 5];
        float base[VEC], zmax[VEC], vsum[VEC], vsq[VEC];
        int kbest[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            // n_i enters through (n_j - n_i) and through the last three channels
            base[v] = (w[v][4] - w[v][1]) * a0 + (w[v][5] - w[v][2]) * a1 + (w[v][6] - w[v][3]) * a2;
            zmax[v] = -CUDART_INF_F; vsum[v] = 0.f; vsq[v] = 0.f; kbest[v] = 0;
        }
        const int32_t *ip = a.idx + ((size_t)b * a.N + i) * k;
        for (int k0 = 0; k0 < k; k0 += 64) {
            const int kn = min(64, k - k0);
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int kk = lane + 32 * t;
                if (kk < kn) {
                    const int j = ip[k0 + kk];
                    const float b0 = xb[(size_t)j * a.ldx + 3], b1 = xb[(size_t)j * a.ldx + 4], b2 = xb[(size_t)j * a.ldx + 5];
                    const float ang = normal_angle(a0, a1, a2, b0, b1, b2);
                    s_edge[warp][kk] = make_float4(ang, b0, b1, b2);
                    const float e[kNF] = {ang, b0 - a0, b1 - a1, b2 - a2, a0, a1, a2};
                    int m = kNF;
#pragma unroll
                    for (int f = 0; f < kNF; ++f) {
                        macc[f] += e[f];
#pragma unroll
                        for (int g = f; g < kNF; ++g) { macc[m] = fmaf(e[f], e[g], macc[m]); ++m; }
                    }
                }
            }
            __syncwarp();
            for (int kk = 0; kk < kn; ++kk) {
                const float4 nb = s_edge[warp][kk];
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    float y = fmaf(w[v][0], nb.x, base[v]);
                    y = fmaf(w[v][1], nb.y, y);
                    y = fmaf(w[v][2], nb.z, y);
                    y = fmaf(w[v][3], nb.w, y);
                    const float z = sg[v] * y;
                    if (z > zmax[v]) { zmax[v] = z; kbest[v] = k0 + kk; }
                    vsum[v] += y;
                    vsq[v] = fmaf(y, y, vsq[v]);
                }
            }
            __syncwarp();
        }
        const size_t o = ((size_t)b * a.N + i) * Cout + c0;
        float ys[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            ys[v] = sg[v] * zmax[v];
            a.arg[o + v] = (unsigned char)kbest[v];
            s1 += (double)vsum[v];
            s2 += (double)vsq[v];
        }
        VecIO<VEC>::st(a.ysel + o, ys);
    }
    red[threadIdx.x][0] = s1;
    red[threadIdx.x][1] = s2;
#pragma unroll
    for (int m = 0; m < kNMom; ++m) {
        double v = (double)macc[m];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULLM, v, o);
        if (lane == 0) s_mom[warp][m] = v;
    }
    __syncthreads();
    if (threadIdx.x < a.G * 2) {
        const int g = threadIdx.x >> 1, which = threadIdx.x & 1;
        const int lanes_per_group = 32 / a.G;
        double s = 0.0;
        for (int wv = 0; wv < kGWarps; ++wv)
            for (int l = g * lanes_per_group; l < (g + 1) * lanes_per_group; ++l) s += red[wv * 32 + l][which];
        a.part[(((size_t)b * gridDim.x + blockIdx.x) * a.G + g) * 2 + which] = s;
    }
    if (threadIdx.x >= 64 && threadIdx.x < 64 + kNMom) {
        const int item = threadIdx.x - 64;
        double s = 0.0;
        for (int wv = 0; wv < kGWarps; ++wv) s += s_mom[wv][item];
        a.mom_part[((size_t)b * gridDim.x + blockIdx.x) * kNMom + item] = s;
    }
}

// mom[b][0:7] = E1, mom[b][7:56] = E2 as a full symmetric 7x7
__global__ void normal_edge_moments_kernel(const double *__restrict__ mom_part, float *__restrict__ mom, int nblk) {
    __shared__ double tot[kNMom];
    const int b = blockIdx.x;
    if (threadIdx.x < kNMom) {
        double s = 0.0;
        for (int i = 0; i < nblk; ++i) s += mom_part[((size_t)b * nblk + i) * kNMom + threadIdx.x];
        tot[threadIdx.x] = s;
    }
    __syncthreads();
    if (threadIdx.x < kNF) mom[b * 56 + threadIdx.x] = (float)tot[threadIdx.x];
    if (threadIdx.x < kNF * kNF) {
        int f = threadIdx.x / kNF, g = threadIdx.x % kNF;
        int lo = f < g ? f : g, hi = f < g ? g : f;
        int item = kNF;
        for (int r = 0; r < lo; ++r) item += kNF - r;
        item += hi - lo;
        mom[b * 56 + kNF + threadIdx.x] = (float)tot[item];
    }
}

struct NBwdArgs {
    const float *x_nc;
    const int32_t *idx;
    const float *ysel, *stats, *gamma, *beta, *gout;
    const unsigned char *arg;
    float *dw_part;        // [B][nblk][Cout][7]
    int N, ldx, Cout, k, G;
    float slope;
};

template <int VEC>
__global__ void __launch_bounds__(kGWarps * 32) normal_edge_bwd_kernel(NBwdArgs a) {
    __shared__ float red[kGWarps][32 * VEC][kNF];
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Cout = a.Cout, c0 = lane * VEC;
    const int cpg = Cout / a.G;
    const float mean = a.stats[((size_t)b * a.G + c0 / cpg) * 2 + 0], rstd = a.stats[((size_t)b * a.G + c0 / cpg) * 2 + 1];
    const float *xb = a.x_nc + (size_t)b * a.N * a.ldx;
    float gm[VEC], bt[VEC], acc[VEC][kNF];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        gm[v] = a.gamma[c0 + v];
        bt[v] = a.beta[c0 + v];
#pragma unroll
        for (int f = 0; f < kNF; ++f) acc[v][f] = 0.f;
    }
    for (int pi = 0; pi < kPtsPerWarp; ++pi) {
        const int i = blockIdx.x * kPtsPerCta + warp * kPtsPerWarp + pi;
        if (i >= a.N) break;
        const float a0 = xb[(size_t)i * a.ldx + 3], a1 = xb[(size_t)i * a.ldx + 4], a2 = xb[(size_t)i * a.ldx + 5];
        const size_t o = ((size_t)b * a.N + i) * Cout + c0;
        const int32_t *ip = a.idx + ((size_t)b * a.N + i) * a.k;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const float yh = (a.ysel[o + v] - mean) * rstd;
            const float u = yh * gm[v] + bt[v];
            const float g = a.gout[o + v];
            const float s = rstd * gm[v] * (u > 0.f ? g : g * a.slope);
            const int j = ip[a.arg[o + v]];
            const float b0 = xb[(size_t)j * a.ldx + 3], b1 = xb[(size_t)j * a.ldx + 4], b2 = xb[(size_t)j * a.ldx + 5];
            const float e[kNF] = {normal_angle(a0, a1, a2, b0, b1, b2), b0 - a0, b1 - a1, b2 - a2, a0, a1, a2};
#pragma unroll
            for (int f = 0; f < kNF; ++f) acc[v][f] = fmaf(s, e[f], acc[v][f]);
        }
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v)
#pragma unroll
        for (int f = 0; f < kNF; ++f) red[warp][c0 + v][f] = acc[v][f];
    __syncthreads();
    for (int e = threadIdx.x; e < Cout * kNF; e += blockDim.x) {
        const int c = e / kNF, f = e % kNF;
        float s = 0.f;
#pragma unroll
        for (int wv = 0; wv < kGWarps; ++wv) s += red[wv][c][f];
        a.dw_part[((size_t)b * gridDim.x + blockIdx.x) * Cout * kNF + e] = s;
    }
}

// dW[c][f] = sum_{b,blk} dw_part + sum_b (A_g E1_b[f] + K_g sum_f' W[c][f'] E2_b[f'][f])
// block (32 entries, 32 slices of the partial list): the partials are summed slice by slice, then across the slices, in a
// fixed order
__global__ void __launch_bounds__(1024) normal_edge_dw_kernel(const float *__restrict__ dw_part, const float *__restrict__ coef,
                                                              const float *__restrict__ mom, const float *__restrict__ weight,
                                                              float *__restrict__ dW, int B, int nblk, int Cout, int G) {
    __shared__ double red[32][33];
    const int e = blockIdx.x * 32 + threadIdx.x;
    const bool live = e < Cout * kNF;
    double s = 0.0;
    if (live)
        for (int i = threadIdx.y; i < B * nblk; i += 32) s += (double)dw_part[(size_t)i * Cout * kNF + e];
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y != 0 || !live) return;
    s = 0.0;
    for (int r = 0; r < 32; ++r) s += red[r][threadIdx.x];
    const int c = e / kNF, f = e % kNF;
    const int g = c / (Cout / G);
    for (int b = 0; b < B; ++b) {
        const float Ag = coef[((size_t)b * G + g) * 2 + 0], Kg = coef[((size_t)b * G + g) * 2 + 1];
        const float *m = mom + b * 56;
        double we2 = 0.0;
        for (int fp = 0; fp < kNF; ++fp) we2 += (double)weight[(size_t)c * kNF + fp] * (double)m[kNF + fp * kNF + f];
        s += (double)Ag * (double)m[f] + (double)Kg * we2;
    }
    dW[e] = (float)s;
}

struct NSaved {
    float *ysel, *stats, *mom;
    unsigned char *arg;
};

static size_t plan_nsaved(const gcanet_normal_edge_desc *d, void *base, NSaved *s) {
    Carver cv(base);
    size_t bn = (size_t)d->B * d->N;
    float *ysel = cv.take<float>(bn * d->Cout);
    unsigned char *arg = cv.take<unsigned char>(bn * d->Cout);
    float *stats = cv.take<float>((size_t)d->B * d->groups * 2);
    float *mom = cv.take<float>((size_t)d->B * 56);
    if (s) { s->ysel = ysel; s->arg = arg; s->stats = stats; s->mom = mom; }
    return cv.off;
}

struct NWs {
    double *part, *mom_part, *sbc;
    float *rpart, *coef, *dw_part;
};

static size_t plan_nws(const gcanet_normal_edge_desc *d, void *base, NWs *w) {
    Carver cv(base);
    int nblk = ceil_div(d->N, kPtsPerCta);
    double *part = cv.take<double>((size_t)d->B * nblk * d->groups * 2);
    double *mom_part = cv.take<double>((size_t)d->B * nblk * kNMom);
    double *sbc = cv.take<double>((size_t)d->B * d->Cout * 2);
    float *rpart = cv.take<float>((size_t)d->B * nblk * d->Cout * 2);
    float *coef = cv.take<float>((size_t)d->B * d->groups * 2);
    float *dw_part = cv.take<float>((size_t)d->B * nblk * d->Cout * kNF);
    if (w) { w->part = part; w->mom_part = mom_part; w->sbc = sbc; w->rpart = rpart; w->coef = coef; w->dw_part = dw_part; }
    return cv.off;
}

static int check_ndesc(const gcanet_normal_edge_desc *d) {
    GCANET_REQUIRE(d != nullptr, "normal_edgeconv: null descriptor");
    GCANET_REQUIRE(d->B >= 1 && d->B <= 65535 && d->N >= 1 && d->k >= 1 && d->k <= 255, "normal_edgeconv: bad shape B=%d N=%d k=%d", d->B, d->N, d->k);
    GCANET_REQUIRE(d->ldx >= 6, "normal_edgeconv: x_nc needs xyz + normals (ldx=%d < 6)", d->ldx);
    GCANET_REQUIRE(d->Cout == 32 || d->Cout == 64 || d->Cout == 128, "normal_edgeconv: Cout=%d must be 32, 64 or 128", d->Cout);
    GCANET_REQUIRE(d->groups >= 1 && d->Cout % d->groups == 0 && 32 % d->groups == 0, "normal_edgeconv: groups=%d must divide 32 and Cout", d->groups);
    GCANET_REQUIRE(d->eps > 0.f, "normal_edgeconv: eps must be positive");
    GCANET_REQUIRE(d->slope >= 0.f, "normal_edgeconv: negative_slope=%g < 0", (double)d->slope);
    GCANET_REQUIRE((long long)d->B * d->N < 2147483647ll, "normal_edgeconv: B * N = %lld does not fit 32-bit row counts", (long long)d->B * d->N);
    return GCANET_OK;
}

template <int VEC>
static int run_nforward(const gcanet_normal_edge_desc *d, const float *x_nc, const int32_t *idx, const float *weight,
                        const float *gamma, const float *beta, float *out_nc, float *out_cn, const NSaved &sv,
                        const NWs &w, cudaStream_t st) {
    const int nblk = ceil_div(d->N, kPtsPerCta);
    NFwdArgs fa{x_nc, idx, weight, gamma, sv.ysel, sv.arg, w.part, w.mom_part, d->N, d->ldx, d->Cout, d->k, d->groups};
    normal_edge_forward_kernel<VEC><<<dim3(nblk, d->B), kGWarps * 32, 0, st>>>(fa);
    GCANET_LAUNCH_OK("normal_edge_forward_kernel");
    double count = (double)(d->Cout / d->groups) * d->N * d->k;
    gn_stats_kernel<<<d->B, 32 * d->groups, 0, st>>>(w.part, sv.stats, nblk, d->groups, count, d->eps);
    GCANET_LAUNCH_OK("gn_stats_kernel");
    normal_edge_moments_kernel<<<d->B, 64, 0, st>>>(w.mom_part, sv.mom, nblk);
    GCANET_LAUNCH_OK("normal_edge_moments_kernel");
    dim3 fg(ceil_div(d->N, 32), d->Cout / 32, d->B), fb(32, 8);
    edge_finish_kernel<<<fg, fb, 0, st>>>(sv.ysel, sv.stats, gamma, beta, out_nc, out_cn, d->N, d->Cout, d->groups, d->slope);
    GCANET_LAUNCH_OK("edge_finish_kernel");
    return GCANET_OK;
}

template <int VEC>
static int run_nbackward(const gcanet_normal_edge_desc *d, const float *x_nc, const int32_t *idx, const float *weight,
                         const float *gamma, const float *beta, const float *gout, const NSaved &sv, float *grad_weight,
                         float *grad_gamma, float *grad_beta, const NWs &w, cudaStream_t st) {
    const int nblk = ceil_div(d->N, kPtsPerCta);
    const int Cout = d->Cout;
    BwdArgs ba{nullptr, sv.ysel, nullptr, sv.stats, gamma, beta, gout, sv.arg, idx, w.rpart, w.coef, nullptr, nullptr,
               d->N, Cout, d->k, d->groups, d->slope};
    edge_bwd_reduce_kernel<VEC><<<dim3(nblk, d->B), kGWarps * 32, 0, st>>>(ba);
    GCANET_LAUNCH_OK("edge_bwd_reduce_kernel");
    double count = (double)(Cout / d->groups) * d->N * d->k;
    edge_bwd_chansum_kernel<<<dim3(ceil_div(Cout, 8), d->B), 256, 0, st>>>(w.rpart, w.sbc, nblk, Cout);
    GCANET_LAUNCH_OK("edge_bwd_chansum_kernel");
    edge_bwd_coef_kernel<<<d->B, 32 * d->groups, 0, st>>>(w.sbc, sv.stats, gamma, w.coef, Cout, d->groups, count);
    GCANET_LAUNCH_OK("edge_bwd_coef_kernel");
    edge_bwd_affine_kernel<<<ceil_div(Cout, 128), 128, 0, st>>>(w.sbc, grad_gamma, grad_beta, d->B, Cout);
    GCANET_LAUNCH_OK("edge_bwd_affine_kernel");
    NBwdArgs na{x_nc, idx, sv.ysel, sv.stats, gamma, beta, gout, sv.arg, w.dw_part, d->N, d->ldx, Cout, d->k, d->groups, d->slope};
    normal_edge_bwd_kernel<VEC><<<dim3(nblk, d->B), kGWarps * 32, 0, st>>>(na);
    GCANET_LAUNCH_OK("normal_edge_bwd_kernel");
    normal_edge_dw_kernel<<<ceil_div(Cout * kNF, 32), dim3(32, 32), 0, st>>>(w.dw_part, w.coef, sv.mom, weight, grad_weight, d->B,
                                                                     nblk, Cout, d->groups);
    GCANET_LAUNCH_OK("normal_edge_dw_kernel");
    return GCANET_OK;
}

}  // namespace gcanet

using namespace gcanet;

extern "C" size_t gcanet_edgeconv_saved_bytes(const gcanet_edgeconv_desc *d) {
    if (check_desc(d) != GCANET_OK) return 0;
    return plan_saved(d, nullptr, nullptr);
}

extern "C" size_t gcanet_edgeconv_workspace_bytes(const gcanet_edgeconv_desc *d) {
    if (check_desc(d) != GCANET_OK) return 0;
    size_t f = plan_fwd(d, nullptr, nullptr), b = plan_bwd(d, nullptr, nullptr);
    return f > b ? f : b;
}

static int check_ws(const char *who, const void *p, size_t have, size_t need) {
    if (p == nullptr || have < need || reinterpret_cast<uintptr_t>(p) % kAlign) {
        set_error("%s: workspace too small or misaligned (%zu given, %zu needed)", who, have, need);
        return GCANET_ERR_WORKSPACE;
    }
    return GCANET_OK;
}

extern "C" int gcanet_edgeconv_forward(const gcanet_edgeconv_desc *d, const float *x_nc, const int32_t *idx,
                                       const float *weight, const float *gamma, const float *beta, float *out_nc,
                                       float *out_cn, void *saved, void *ws, size_t ws_bytes, gcanet_stream_t stream) {
    int rc = check_desc(d);
    if (rc) return rc;
    GCANET_REQUIRE(x_nc && idx && weight && gamma && beta && out_nc && saved, "edgeconv_forward: null pointer");
    GCANET_REQUIRE(reinterpret_cast<uintptr_t>(saved) % kAlign == 0, "edgeconv_forward: saved buffer must be 256-byte aligned");
    rc = check_ws("edgeconv_forward", ws, ws_bytes, plan_fwd(d, nullptr, nullptr));
    if (rc) return rc;
    Saved sv; FwdWs w;
    plan_saved(d, saved, &sv);
    plan_fwd(d, ws, &w);
    cudaStream_t st = as_stream(stream);
    switch (d->Cout / 32) {
        case 1: return run_forward<1>(d, x_nc, idx, weight, gamma, beta, out_nc, out_cn, sv, w, st);
        case 2: return run_forward<2>(d, x_nc, idx, weight, gamma, beta, out_nc, out_cn, sv, w, st);
        case 4: return run_forward<4>(d, x_nc, idx, weight, gamma, beta, out_nc, out_cn, sv, w, st);
        default: return run_forward<8>(d, x_nc, idx, weight, gamma, beta, out_nc, out_cn, sv, w, st);
    }
}

extern "C" int gcanet_edgeconv_backward(const gcanet_edgeconv_desc *d, const float *x_nc, const int32_t *idx,
                                        const float *weight, const float *gamma, const float *beta,
                                        const float *grad_out_nc, const void *saved, float *grad_x_nc,
                                        float *grad_weight, float *grad_gamma, float *grad_beta, void *ws,
                                        size_t ws_bytes, gcanet_stream_t stream) {
    int rc = check_desc(d);
    if (rc) return rc;
    GCANET_REQUIRE(x_nc && idx && weight && gamma && beta && grad_out_nc && saved && grad_weight && grad_gamma && grad_beta,
                   "edgeconv_backward: null pointer");
    rc = check_ws("edgeconv_backward", ws, ws_bytes, plan_bwd(d, nullptr, nullptr));
    if (rc) return rc;
    Saved sv; BwdWs w;
    plan_saved(d, const_cast<void *>(saved), &sv);
    plan_bwd(d, ws, &w);
    cudaStream_t st = as_stream(stream);
    switch (d->Cout / 32) {
        case 1: return run_backward<1>(d, x_nc, idx, weight, gamma, beta, grad_out_nc, sv, grad_x_nc, grad_weight, grad_gamma, grad_beta, w, st);
        case 2: return run_backward<2>(d, x_nc, idx, weight, gamma, beta, grad_out_nc, sv, grad_x_nc, grad_weight, grad_gamma, grad_beta, w, st);
        case 4: return run_backward<4>(d, x_nc, idx, weight, gamma, beta, grad_out_nc, sv, grad_x_nc, grad_weight, grad_gamma, grad_beta, w, st);
        default: return run_backward<8>(d, x_nc, idx, weight, gamma, beta, grad_out_nc, sv, grad_x_nc, grad_weight, grad_gamma, grad_beta, w, st);
    }
}

extern "C" size_t gcanet_normal_edgeconv_saved_bytes(const gcanet_normal_edge_desc *d) {
    if (check_ndesc(d) != GCANET_OK) return 0;
    return plan_nsaved(d, nullptr, nullptr);
}

extern "C" size_t gcanet_normal_edgeconv_workspace_bytes(const gcanet_normal_edge_desc *d) {
    if (check_ndesc(d) != GCANET_OK) return 0;
    return plan_nws(d, nullptr, nullptr);
}

extern "C" int gcanet_normal_edgeconv_forward(const gcanet_normal_edge_desc *d, const float *x_nc, const int32_t *idx,
                                              const float *weight, const float *gamma, const float *beta,
                                              float *out_nc, float *out_cn, void *saved, void *ws, size_t ws_bytes,
                                              gcanet_stream_t stream) {
    int rc = check_ndesc(d);
    if (rc) return rc;
    GCANET_REQUIRE(x_nc && idx && weight && gamma && beta && out_nc && saved, "normal_edgeconv_forward: null pointer");
    GCANET_REQUIRE(reinterpret_cast<uintptr_t>(saved) % kAlign == 0, "normal_edgeconv_forward: saved buffer must be 256-byte aligned");
    rc = check_ws("normal_edgeconv_forward", ws, ws_bytes, plan_nws(d, nullptr, nullptr));
    if (rc) return rc;
    NSaved sv; NWs w;
    plan_nsaved(d, saved, &sv);
    plan_nws(d, ws, &w);
    cudaStream_t st = as_stream(stream);
    switch (d->Cout / 32) {
        case 1: return run_nforward<1>(d, x_nc, idx, weight, gamma, beta, out_nc, out_cn, sv, w, st);
        case 2: return run_nforward<2>(d, x_nc, idx, weight, gamma, beta, out_nc, out_cn, sv, w, st);
        default: return run_nforward<4>(d, x_nc, idx, weight, gamma, beta, out_nc, out_cn, sv, w, st);
    }
}

extern "C" int gcanet_normal_edgeconv_backward(const gcanet_normal_edge_desc *d, const float *x_nc, const int32_t *idx,
                                               const float *weight, const float *gamma, const float *beta,
                                               const float *grad_out_nc, const void *saved, float *grad_weight,
                                               float *grad_gamma, float *grad_beta, void *ws, size_t ws_bytes,
                                               gcanet_stream_t stream) {
    int rc = check_ndesc(d);
    if (rc) return rc;
    GCANET_REQUIRE(x_nc && idx && weight && gamma && beta && grad_out_nc && saved && grad_weight && grad_gamma && grad_beta,
                   "normal_edgeconv_backward: null pointer");
    rc = check_ws("normal_edgeconv_backward", ws, ws_bytes, plan_nws(d, nullptr, nullptr));
    if (rc) return rc;
    NSaved sv; NWs w;
    plan_nsaved(d, const_cast<void *>(saved), &sv);
    plan_nws(d, ws, &w);
    cudaStream_t st = as_stream(stream);
    switch (d->Cout / 32) {
        case 1: return run_nbackward<1>(d, x_nc, idx, weight, gamma, beta, grad_out_nc, sv, grad_weight, grad_gamma, grad_beta, w, st);
        case 2: return run_nbackward<2>(d, x_nc, idx, weight, gamma, beta, grad_out_nc, sv, grad_weight, grad_gamma, grad_beta, w, st);
        default: return run_nbackward<4>(d, x_nc, idx, weight, gamma, beta, grad_out_nc, sv, grad_weight, grad_gamma, grad_beta, w, st);
    }
}
