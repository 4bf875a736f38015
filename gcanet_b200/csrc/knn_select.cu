// CUDA-core brute-force kNN with a fused streaming top-k (no N x N matrix in memory).
//
// One warp owns R queries; the 32 lanes each take one reference point of the current
// group, so a group of 32 candidates costs C*R FMAs per lane plus one compare per query.
// Every query keeps a candidate list in shared memory and a threshold (an upper bound of
// its current k-th distance): lanes whose candidate beats the threshold append it with one
// ballot + prefix-popcount (no per-candidate serialisation).  A list that is about to
// overflow is shrunk by its warp: bisection for a bound with count(d <= bound) >= k, keep
// d <= bound, the bound becomes the new threshold.  After the scan the survivors are ranked
// by (distance, index) and the k best written in order.  Only O(k log(N/k)) candidates per
// query ever reach a list, so the scan is FMA/compare bound.
//
// Arithmetic follows the reference so that index sets agree up to fp32 ties:
//   L2   (M4:36-38)  d = fl(fl(|x_j|^2 - 2 x_i.x_j) + |x_i|^2), norms summed without FMA
//   PN   (M4:61-73)  d = d_p * (1 + (2 - 2 n_i.n_j)),  d_p as above on channels 0..2
//   SSD  (KNN/csrc/cuda/knn.cu:73-76)  d = sum_c fma(t, t, d), t = ref_c - query_c
// Ties: ranking by (distance, index) keeps the lower reference index first and a later
// candidate equal to the k-th never displaces it (knn.cu:125,149).
#include "knn_lists.cuh"

namespace gcanet {

enum { METRIC_L2 = 0, METRIC_PN = 1, METRIC_SSD = 2 };

constexpr int kWarps = 8;               // warps per CTA
constexpr int kThreads = kWarps * 32;

// ---------------------------------------------------------------------------------
// squared norms, reference order: sum_c fl(x_c * x_c), sequential, no FMA contraction
// ---------------------------------------------------------------------------------
__global__ void sqnorm_kernel(const float *__restrict__ x, float *__restrict__ out, int C, int Cuse, int N) {
    int b = blockIdx.y;
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float *p = x + (size_t)b * C * N + n;
    float s = 0.f;
    for (int c = 0; c < Cuse; ++c) {
        float v = p[(size_t)c * N];
        s = __fadd_rn(s, __fmul_rn(v, v));
    }
    out[(size_t)b * N + n] = s;
}

struct ScanArgs {
    const float *ref;       // [B][C][Nr]
    const float *qry;       // [B][C][Nq]
    const float *ref_norm;  // [B][Nr]   (L2 / PN)
    const float *qry_norm;  // [B][Nq]
    int C, Nr, Nq, k;
    int step, kout;         // dilation: list position p is written to column p/step when p % step == 0
    int64_t *idx64;
    int32_t *idx32;
    float *dist;            // optional; sqrt(d) is written (KNN_CUDA path)
    int k_major;            // 0: out[b][q][col]   1: out[b][col][q]
    int index_base;
    int TR;                 // reference tile (multiple of 32)
    const int *qlist;       // optional [B][Nq] + qcount [B]: only the listed queries are computed / written (the
    const int *qcount;      //   tensor-core path's overflow rows); CTAs then stride over the list
    int list_min;           // list mode: clouds with fewer listed rows than this are left alone (0 = no limit)
    const int *cloud_filter;  // optional [B]: only clouds with a non-zero flag are computed
};

// CDIM > 0: compile-time dimension, queries in registers.  CDIM == 0: runtime C, queries in smem.
// SL: candidate-list capacity in units of 32 entries (needs k + kSlack + 32 <= 32 * SL).
template <int CDIM, int METRIC, int SL, int R>
__global__ void __launch_bounds__(kThreads) knn_scan_kernel(ScanArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int C = CDIM > 0 ? CDIM : a.C;
    const int TR = a.TR;
    constexpr int QPB = kWarps * R;
    constexpr int CAP = 32 * SL;
    float *rs = smem;                 // [C][TR]
    float *rn = rs + (size_t)C * TR;  // [TR]
    float *qs = rn + TR;              // [C][QPB]
    float *lds = qs + (size_t)C * QPB;                       // [QPB][CAP] candidate distances
    int *lis = reinterpret_cast<int *>(lds + QPB * CAP);     // [QPB][CAP] candidate indices

    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float *ref = a.ref + (size_t)b * C * a.Nr;
    const float *qry = a.qry + (size_t)b * C * a.Nq;

    if (a.cloud_filter != nullptr && a.cloud_filter[b] == 0) return;
    // list mode: slot s of this cloud's work is query qlist[b][s]; otherwise slot = query
    const int *ql = a.qlist ? a.qlist + (size_t)b * a.Nq : nullptr;
    const int nq = ql ? a.qcount[b] : a.Nq;
    if (ql && a.list_min > 0 && nq < a.list_min) return;      // few listed rows: knn_row_kernel took them
    for (int vbx = blockIdx.x; vbx * QPB < nq; vbx += gridDim.x) {
    const int q0 = vbx * QPB + warp * R;
    __syncthreads();   // shared memory of the previous round is free

    // stage this CTA's queries (zero for out-of-range ones)
    for (int e = threadIdx.x; e < C * QPB; e += kThreads) {
        int c = e / QPB, qq = e % QPB;
        int sl = vbx * QPB + qq;
        qs[e] = sl < nq ? qry[(size_t)c * a.Nq + (ql ? ql[sl] : sl)] : 0.f;
    }
    __syncthreads();

    float qreg[R][CDIM > 0 ? CDIM : 1];
    float qn[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (CDIM > 0) {
#pragma unroll
            for (int c = 0; c < (CDIM > 0 ? CDIM : 1); ++c) qreg[r][c] = qs[c * QPB + warp * R + r];
        }
        qn[r] = 0.f;
        if (METRIC != METRIC_SSD) {
            int sl = q0 + r;
            qn[r] = sl < nq ? a.qry_norm[(size_t)b * a.Nq + (ql ? ql[sl] : sl)] : 0.f;
        }
    }

    float thr[R];
    int cnt[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { thr[r] = CUDART_INF_F; cnt[r] = 0; }
    float *my_ld = lds + (size_t)(warp * R) * CAP;
    int *my_li = lis + (size_t)(warp * R) * CAP;

    for (int t0 = 0; t0 < a.Nr; t0 += TR) {
        __syncthreads();   // previous tile fully consumed
        for (int e = threadIdx.x; e < C * TR; e += kThreads) {
            int c = e / TR, jj = e % TR;
            int j = t0 + jj;
            rs[e] = j < a.Nr ? ref[(size_t)c * a.Nr + j] : 0.f;
        }
        if (METRIC != METRIC_SSD) {
            for (int jj = threadIdx.x; jj < TR; jj += kThreads) {
                int j = t0 + jj;
                rn[jj] = j < a.Nr ? a.ref_norm[(size_t)b * a.Nr + j] : 0.f;
            }
        }
        __syncthreads();

        const int groups = min(TR, a.Nr - t0 + 31) >> 5;
        for (int g = 0; g < groups; ++g) {
            const int jj = g * 32 + lane;
            const int j = t0 + jj;
            float d[R];
            if (CDIM > 0) {
                float rv[CDIM > 0 ? CDIM : 1];
#pragma unroll
                for (int c = 0; c < (CDIM > 0 ? CDIM : 1); ++c) rv[c] = rs[c * TR + jj];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (METRIC == METRIC_SSD) {
                        float s = 0.f;
#pragma unroll
                        for (int c = 0; c < (CDIM > 0 ? CDIM : 1); ++c) {
                            float t = rv[c] - qreg[r][c];
                            s = fmaf(t, t, s);
                        }
                        d[r] = s;
                    } else if (METRIC == METRIC_L2) {
                        float t = __fmul_rn(qreg[r][0], rv[0]);
#pragma unroll
                        for (int c = 1; c < (CDIM > 0 ? CDIM : 1); ++c) t = fmaf(qreg[r][c], rv[c], t);
                        d[r] = __fadd_rn(fmaf(-2.f, t, rn[jj]), qn[r]);
                    } else {   // METRIC_PN, CDIM == 6
                        float tp = __fmul_rn(qreg[r][0], rv[0]);
                        tp = fmaf(qreg[r][1 % (CDIM > 0 ? CDIM : 1)], rv[1 % (CDIM > 0 ? CDIM : 1)], tp);
                        tp = fmaf(qreg[r][2 % (CDIM > 0 ? CDIM : 1)], rv[2 % (CDIM > 0 ? CDIM : 1)], tp);
                        float tn = __fmul_rn(qreg[r][3 % (CDIM > 0 ? CDIM : 1)], rv[3 % (CDIM > 0 ? CDIM : 1)]);
                        tn = fmaf(qreg[r][4 % (CDIM > 0 ? CDIM : 1)], rv[4 % (CDIM > 0 ? CDIM : 1)], tn);
                        tn = fmaf(qreg[r][5 % (CDIM > 0 ? CDIM : 1)], rv[5 % (CDIM > 0 ? CDIM : 1)], tn);
                        float dp = __fadd_rn(fmaf(-2.f, tp, rn[jj]), qn[r]);
                        float dn = fmaf(-2.f, tn, 2.f);
                        d[r] = __fmul_rn(dp, __fadd_rn(1.f, dn));
                    }
                }
            } else {
                float acc[R];
#pragma unroll
                for (int r = 0; r < R; ++r) acc[r] = 0.f;
                const float *rp = rs + jj;
                const float *qp = qs + warp * R;
#pragma unroll 4
                for (int c = 0; c < C; ++c) {
                    float rvv = rp[(size_t)c * TR];
                    float qv[R];
                    if (R % 4 == 0) {
#pragma unroll
                        for (int r4 = 0; r4 < R / 4; ++r4) {
                            float4 q4 = *reinterpret_cast<const float4 *>(qp + c * QPB + r4 * 4);
                            qv[r4 * 4 + 0] = q4.x; qv[r4 * 4 + 1] = q4.y; qv[r4 * 4 + 2] = q4.z; qv[r4 * 4 + 3] = q4.w;
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < R; ++r) qv[r] = qp[c * QPB + r];
                    }
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        if (METRIC == METRIC_SSD) {
                            float t = rvv - qv[r];
                            acc[r] = fmaf(t, t, acc[r]);
                        } else {
                            acc[r] = fmaf(qv[r], rvv, acc[r]);
                        }
                    }
                }
#pragma unroll
                for (int r = 0; r < R; ++r)
                    d[r] = METRIC == METRIC_SSD ? acc[r] : __fadd_rn(fmaf(-2.f, acc[r], rn[jj]), qn[r]);
            }
            const bool valid = j < a.Nr;
            static_for<0, R>([&](auto rc) {
                constexpr int r = decltype(rc)::value;
                const bool pass = valid && d[r] < thr[r];
                const unsigned m = __ballot_sync(FULL, pass);
                if (m) {
                    if (pass) {
                        const int p = cnt[r] + __popc(m & ((1u << lane) - 1));
                        my_ld[r * CAP + p] = d[r];
                        my_li[r * CAP + p] = j;
                    }
                    cnt[r] += __popc(m);
                    if (cnt[r] > CAP - 32) {
                        __syncwarp();
                        float t = thr[r];      // by-reference argument of a non-inlined call: keep thr[] itself in registers
                        cnt[r] = shrink_list<SL>(my_ld + r * CAP, my_li + r * CAP, cnt[r], a.k, CAP - 32, lane, t);
                        thr[r] = t;
                    }
                }
            });
        }
    }

    // rank the survivors and write the k best in order
    __syncwarp();
    static_for<0, R>([&](auto rc) {
        constexpr int r = decltype(rc)::value;
        if (q0 + r >= nq) return;
        float *ld = my_ld + r * CAP;
        int *li = my_li + r * CAP;
        float t = thr[r];
        int n = cnt[r];
        if (n > a.k + kSlack) n = shrink_list<SL>(ld, li, n, a.k, CAP, lane, t);   // cheap pre-shrink
        rank_cut<SL>(ld, li, n, a.k, lane);                                         // k best, in order
    });
    __syncwarp();
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (q0 + r >= nq) continue;
        const int q = ql ? ql[q0 + r] : q0 + r;
        for (int p = lane; p < a.k; p += 32) {
            if (p % a.step) continue;
            const int col = p / a.step;
            const size_t o = a.k_major ? ((size_t)b * a.kout + col) * a.Nq + q
                                       : ((size_t)b * a.Nq + q) * a.kout + col;
            const int id = my_li[r * CAP + p] + a.index_base;
            if (a.idx64) a.idx64[o] = id;
            if (a.idx32) a.idx32[o] = id;
            if (a.dist) a.dist[o] = sqrtf(my_ld[r * CAP + p]);
        }
    }
    }   // slots of this CTA
}

// ---------------------------------------------------------------------------------
// k > 128: one query per warp, the list lives in shared memory.  Generic and slow;
// exists so the KNN_CUDA entry point covers the reference's own test grid (k = 400).
// ---------------------------------------------------------------------------------
template <int METRIC>
__global__ void __launch_bounds__(kThreads) knn_scan_bigk_kernel(ScanArgs a) {
    extern __shared__ float smem[];
    const int C = a.C, TR = a.TR, k = a.k;
    float *rs = smem;                        // [C][TR]
    float *rn = rs + (size_t)C * TR;         // [TR]
    float *qs = rn + TR;                     // [kWarps][C]
    float *lv = qs + kWarps * C;             // [kWarps][k]
    int *li = reinterpret_cast<int *>(lv + (size_t)kWarps * k);

    const int b = blockIdx.y;
    if (a.cloud_filter != nullptr && a.cloud_filter[b] == 0) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * kWarps + warp;
    const float *ref = a.ref + (size_t)b * C * a.Nr;
    const float *qry = a.qry + (size_t)b * C * a.Nq;
    float *mv = lv + (size_t)warp * k;
    int *mi = li + (size_t)warp * k;

    for (int c = lane; c < C; c += 32) qs[warp * C + c] = q < a.Nq ? qry[(size_t)c * a.Nq + q] : 0.f;
    for (int p = lane; p < k; p += 32) { mv[p] = CUDART_INF_F; mi[p] = 0; }
    float qn = 0.f;
    if (METRIC != METRIC_SSD && q < a.Nq) qn = a.qry_norm[(size_t)b * a.Nq + q];
    float thr = CUDART_INF_F;

    for (int t0 = 0; t0 < a.Nr; t0 += TR) {
        __syncthreads();
        for (int e = threadIdx.x; e < C * TR; e += kThreads) {
            int c = e / TR, jj = e % TR;
            int j = t0 + jj;
            rs[e] = j < a.Nr ? ref[(size_t)c * a.Nr + j] : 0.f;
        }
        if (METRIC != METRIC_SSD)
            for (int jj = threadIdx.x; jj < TR; jj += kThreads) {
                int j = t0 + jj;
                rn[jj] = j < a.Nr ? a.ref_norm[(size_t)b * a.Nr + j] : 0.f;
            }
        __syncthreads();
        const int groups = min(TR, a.Nr - t0 + 31) >> 5;
        for (int g = 0; g < groups; ++g) {
            const int jj = g * 32 + lane;
            float d;
            if (METRIC == METRIC_PN) {
                const float *qq = qs + warp * C;
                float tp = __fmul_rn(qq[0], rs[jj]);
                tp = fmaf(qq[1], rs[TR + jj], tp);
                tp = fmaf(qq[2], rs[2 * TR + jj], tp);
                float tn = __fmul_rn(qq[3], rs[3 * TR + jj]);
                tn = fmaf(qq[4], rs[4 * TR + jj], tn);
                tn = fmaf(qq[5], rs[5 * TR + jj], tn);
                float dp = __fadd_rn(fmaf(-2.f, tp, rn[jj]), qn);
                d = __fmul_rn(dp, __fadd_rn(1.f, fmaf(-2.f, tn, 2.f)));
            } else {
                float acc = 0.f;
                for (int c = 0; c < C; ++c) {
                    float rv = rs[(size_t)c * TR + jj], qv = qs[warp * C + c];
                    if (METRIC == METRIC_SSD) { float t = rv - qv; acc = fmaf(t, t, acc); }
                    else acc = fmaf(qv, rv, acc);
                }
                d = METRIC == METRIC_SSD ? acc : __fadd_rn(fmaf(-2.f, acc, rn[jj]), qn);
            }
            if (t0 + jj >= a.Nr) d = CUDART_INF_F;
            unsigned m = __ballot_sync(FULL, d < thr);
            while (m) {
                int src = __ffs(m) - 1;
                m &= m - 1;
                float dd = __shfl_sync(FULL, d, src);
                if (!(dd < thr)) continue;
                int jnew = t0 + g * 32 + src;
                int pos = 0;
                for (int base = 0; base < k; base += 32) {
                    int p = base + lane;
                    pos += __popc(__ballot_sync(FULL, p < k && mv[p] <= dd));
                }
                for (int base = ((k - 1) >> 5) << 5; base >= 0; base -= 32) {
                    int p = base + lane;
                    float nv = 0.f; int ni = 0; bool w = false;
                    if (p < k && p > pos) { nv = mv[p - 1]; ni = mi[p - 1]; w = true; }
                    else if (p == pos) { nv = dd; ni = jnew; w = true; }
                    __syncwarp();
                    if (w) { mv[p] = nv; mi[p] = ni; }
                    __syncwarp();
                    if (base <= pos) break;
                }
                thr = mv[k - 1];
            }
        }
    }
    __syncwarp();
    if (q < a.Nq) {
        for (int p = lane; p < k; p += 32) {
            if (p % a.step) continue;
            int col = p / a.step;
            size_t o = a.k_major ? ((size_t)b * a.kout + col) * a.Nq + q : ((size_t)b * a.Nq + q) * a.kout + col;
            int id = mi[p] + a.index_base;
            if (a.idx64) a.idx64[o] = id;
            if (a.idx32) a.idx32[o] = id;
            if (a.dist) a.dist[o] = sqrtf(mv[p]);
        }
    }
}

// ---------------------------------------------------------------------------------
// host dispatch
// ---------------------------------------------------------------------------------
static int pick_tile(int C, int Nr) {
    int tr = 16384 / (C > 0 ? C : 1);     // <= 64 KB of reference rows
    tr = tr / 32 * 32;
    if (tr > 1024) tr = 1024;
    if (tr < 32) tr = 32;
    int need = (Nr + 31) / 32 * 32;
    return tr < need ? tr : need;
}

template <int CDIM, int METRIC, int SL, int R>
static int launch_scan(ScanArgs a, int B, cudaStream_t st) {
    constexpr int QPB = kWarps * R;
    size_t smem = ((size_t)a.C * a.TR + a.TR + (size_t)a.C * QPB + 2 * (size_t)QPB * 32 * SL) * sizeof(float);
    auto kern = knn_scan_kernel<CDIM, METRIC, SL, R>;
    if (smem > 48 * 1024) GCANET_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int gx = ceil_div(a.Nq, QPB);
    if (a.qlist != nullptr && gx > 48) gx = 48;      // list mode: a few CTAs per cloud stride over the (usually empty) list
    dim3 grid(gx, B);
    kern<<<grid, kThreads, smem, st>>>(a);
    GCANET_LAUNCH_OK("knn_scan_kernel");
    return GCANET_OK;
}

template <int CDIM, int METRIC, int R>
static int launch_scan_k(ScanArgs a, int B, cudaStream_t st) {
    // list capacity 32*SL holds the k + kSlack survivors of a shrink, one 32-wide batch of new
    // candidates, and >= 48 free slots so that shrinks stay rare
    if (a.k <= 40) return launch_scan<CDIM, METRIC, 4, R>(a, B, st);
    if (a.k <= 104) return launch_scan<CDIM, METRIC, 6, R>(a, B, st);
    return launch_scan<CDIM, METRIC, 8, R>(a, B, st);
}

template <int METRIC>
static int launch_bigk(ScanArgs a, int B, cudaStream_t st) {
    // shrink the reference tile so tile + lists fit in 200 KB
    size_t list_bytes = (size_t)kWarps * a.k * 8 + (size_t)kWarps * a.C * 4;
    int tr = a.TR;
    while (tr > 32 && ((size_t)a.C * tr + tr) * 4 + list_bytes > 200 * 1024) tr -= 32;
    a.TR = tr;
    size_t smem = ((size_t)a.C * tr + tr) * 4 + list_bytes;
    if (smem > 220 * 1024) { set_error("knn: C=%d with k=%d does not fit in shared memory", a.C, a.k); return GCANET_ERR_INVALID_ARGUMENT; }
    auto kern = knn_scan_bigk_kernel<METRIC>;
    GCANET_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(a.Nq, kWarps), B);
    kern<<<grid, kThreads, smem, st>>>(a);
    GCANET_LAUNCH_OK("knn_scan_bigk_kernel");
    return GCANET_OK;
}

static int launch_sqnorm(const float *x, float *out, int B, int C, int Cuse, int N, cudaStream_t st) {
    dim3 grid(ceil_div(N, 256), B);
    sqnorm_kernel<<<grid, 256, 0, st>>>(x, out, C, Cuse, N);
    GCANET_LAUNCH_OK("sqnorm_kernel");
    return GCANET_OK;
}

int launch_sqnorm_public(const float *x, float *out, int B, int C, int Cuse, int N, cudaStream_t st) {
    return launch_sqnorm(x, out, B, C, Cuse, N, st);
}

static int scan_self(const float *x, const float *norms, const int *qlist, const int *qcount, int B, int C, int N, int k1, int k2,
                     int metric, int64_t *idx64, int32_t *idx32, cudaStream_t st, const int *cloud_filter = nullptr,
                     int list_min = 0);

// ---------------------------------------------------------------------------------
// One CTA per listed query (the tensor-core path's overflow rows: massive ties).  The batched scan above needs
// ~1 ms per CTA whatever the number of queries it holds, which made a handful of overflow rows cost more than the
// rest of the call.  Here all N distances of the query go to shared memory (same arithmetic as the scan: fma chain
// over the channels, fl(fl(|x_j|^2 - 2 t) + |x_i|^2)), the k smallest by (distance, index) are found by block-wide
// bisection -- on the distance, then on the index among the points tied at the k-th distance -- and ranked.
// ---------------------------------------------------------------------------------
constexpr int kRowFallbackMax = 512;    // listed rows per cloud up to which the per-row kernel is used

struct RowArgs {
    const float *x;         // [B][C][N]
    const float *norm;      // [B][N]
    const int *qlist;       // [B][N]
    const int *qcount;      // [B]
    int C, N, k, step, kout;
    int64_t *idx64;
    int32_t *idx32;
};

__device__ __forceinline__ int block_sum_256(int v, int *s_red) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    __syncthreads();                                   // s_red free again
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    int t = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_red[w];
    return t;
}

__global__ void __launch_bounds__(256) knn_row_kernel(RowArgs a) {
    extern __shared__ __align__(16) float smem[];
    float *sd = smem;                                   // [N] distances of the current query
    float *s_q = sd + a.N;                              // [C]
    float *s_seld = s_q + a.C;                          // [k]
    int *s_selj = reinterpret_cast<int *>(s_seld + a.k);// [k]
    __shared__ int s_red[8];
    __shared__ float s_fred[2][8];
    __shared__ int s_cnt;
    const int b = blockIdx.y, tid = threadIdx.x;
    const int N = a.N, C = a.C, k = a.k;
    const float *xb = a.x + (size_t)b * C * N;
    const float *nb = a.norm + (size_t)b * N;
    const int nq = a.qcount[b];
    if (nq > kRowFallbackMax) return;                   // many rows: the batched scan (launched next) is cheaper per row
    for (int slot = blockIdx.x; slot < nq; slot += gridDim.x) {
        const int q = a.qlist[(size_t)b * N + slot];
        __syncthreads();
        for (int c = tid; c < C; c += 256) s_q[c] = xb[(size_t)c * N + q];
        if (tid == 0) s_cnt = 0;
        __syncthreads();
        const float qn = nb[q];
        float mn = CUDART_INF_F, mx = -CUDART_INF_F;
        for (int j0 = tid; j0 < N; j0 += 4 * 256) {         // four points per thread and step: four independent fma chains
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            int jj[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) jj[u] = min(j0 + u * 256, N - 1);
#pragma unroll 8
            for (int c = 0; c < C; ++c) {                    // (unrolled: 32 independent loads in flight per thread)
                const float qc = s_q[c];
                const float *xr = xb + (size_t)c * N;
#pragma unroll
                for (int u = 0; u < 4; ++u) acc[u] = fmaf(qc, __ldg(xr + jj[u]), acc[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (j0 + u * 256 < N) {
                    const float d = __fadd_rn(fmaf(-2.f, acc[u], nb[jj[u]]), qn);
                    sd[jj[u]] = d;
                    mn = fminf(mn, d);
                    mx = fmaxf(mx, d);
                }
            }
        }
        for (int o = 16; o; o >>= 1) {
            mn = fminf(mn, __shfl_xor_sync(FULL, mn, o));
            mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
        }
        if ((tid & 31) == 0) { s_fred[0][tid >> 5] = mn; s_fred[1][tid >> 5] = mx; }
        __syncthreads();
        mn = s_fred[0][0]; mx = s_fred[1][0];
#pragma unroll
        for (int w = 1; w < 8; ++w) { mn = fminf(mn, s_fred[0][w]); mx = fmaxf(mx, s_fred[1][w]); }
        // distance bisection: hi keeps count(d <= hi) >= k, lo keeps count(d <= lo) < k (or lo = min)
        float lo = mn, hi = mx;
        int c_hi = N;
        for (int it = 0; it < 80 && c_hi > k; ++it) {
            const float mid = 0.5f * lo + 0.5f * hi;
            if (!(mid > lo && mid < hi)) break;
            int c = 0;
            for (int j = tid; j < N; j += 256) c += (sd[j] <= mid) ? 1 : 0;
            c = block_sum_256(c, s_red);
            if (c >= k) { hi = mid; c_hi = c; } else lo = mid;
        }
        // ties at hi: everything strictly below hi is in (count below), the rest of the k come from the points with
        // d == hi in index order.  (If the loop ended with c_hi == k there is nothing to cut: J = N.)
        int J = N;
        if (c_hi > k) {
            int below = 0;
            for (int j = tid; j < N; j += 256) below += (sd[j] < hi) ? 1 : 0;
            below = block_sum_256(below, s_red);
            if (below >= k) {
                // only when lo was never proven (count(d <= min) >= k): the ties sit at the minimum itself
                hi = lo;
                below = 0;
                for (int j = tid; j < N; j += 256) below += (sd[j] < hi) ? 1 : 0;
                below = block_sum_256(below, s_red);
            }
            const int need = k - below;                     // >= 1
            int jl = -1, jh = N - 1;                        // count(d == hi && j <= jh) >= need
            while (jh - jl > 1) {
                const int jm = (jl + jh) >> 1;
                int c = 0;
                for (int j = tid; j < N; j += 256) c += (sd[j] == hi && j <= jm) ? 1 : 0;
                c = block_sum_256(c, s_red);
                if (c >= need) jh = jm; else jl = jm;
            }
            J = jh;
        }
        __syncthreads();
        for (int j = tid; j < N; j += 256) {
            const float d = sd[j];
            if (d < hi || (d == hi && j <= J)) {
                const int p = atomicAdd(&s_cnt, 1);
                if (p < k) { s_seld[p] = d; s_selj[p] = j; }
            }
        }
        __syncthreads();
        const int m = min(s_cnt, k);
        for (int e = tid; e < m; e += 256) {
            const float d = s_seld[e];
            const int j = s_selj[e];
            int rank = 0;
            for (int f = 0; f < m; ++f) {
                const float od = s_seld[f];
                const int oj = s_selj[f];
                rank += (od < d || (od == d && oj < j)) ? 1 : 0;
            }
            if (rank % a.step == 0) {
                const size_t o = ((size_t)b * N + q) * a.kout + rank / a.step;
                if (a.idx64) a.idx64[o] = j;
                if (a.idx32) a.idx32[o] = j;
            }
        }
    }
}

// re-runs the CUDA-core scan for the queries listed in qlist[b][0 .. qcount[b]) (tensor-core path overflow)
int knn_fallback_rows(const float *x, const float *norms, const int *qlist, const int *qcount, int B, int C, int N, int k1, int k2,
                      int64_t *idx64, int32_t *idx32, cudaStream_t st) {
    const size_t smem = ((size_t)N + C + 2 * (size_t)k2) * sizeof(float);
    if (smem <= 200 * 1024) {
        RowArgs a{x, norms, qlist, qcount, C, N, k2, k2 / k1, gcanet_knn_graph_columns(k1, k2), idx64, idx32};
        if (smem > 48 * 1024) GCANET_CUDA_OK(cudaFuncSetAttribute(knn_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        knn_row_kernel<<<dim3(64, B), 256, smem, st>>>(a);
        GCANET_LAUNCH_OK("knn_row_kernel");
        // clouds with more than kRowFallbackMax listed rows (degenerate inputs) take the batched scan instead
        return scan_self(x, norms, qlist, qcount, B, C, N, k1, k2, GCANET_METRIC_L2, idx64, idx32, st, nullptr, kRowFallbackMax + 1);
    }
    return scan_self(x, norms, qlist, qcount, B, C, N, k1, k2, GCANET_METRIC_L2, idx64, idx32, st);
}

int knn_graph_cuda_cores(const float *x, int B, int C, int N, int k1, int k2, int metric, int64_t *idx64,
                         int32_t *idx32, float *norms, cudaStream_t st) {
    int rc = launch_sqnorm(x, norms, B, C, metric == GCANET_METRIC_POINTS_NORMALS ? 3 : C, N, st);
    if (rc) return rc;
    return scan_self(x, norms, nullptr, nullptr, B, C, N, k1, k2, metric, idx64, idx32, st);
}

static int scan_self(const float *x, const float *norms, const int *qlist, const int *qcount, int B, int C, int N, int k1, int k2,
                     int metric, int64_t *idx64, int32_t *idx32, cudaStream_t st, const int *cloud_filter, int list_min) {
    ScanArgs a{};
    a.qlist = qlist;
    a.qcount = qcount;
    a.list_min = list_min;
    a.cloud_filter = cloud_filter;
    a.ref = x; a.qry = x; a.ref_norm = norms; a.qry_norm = norms;
    a.C = C; a.Nr = N; a.Nq = N; a.k = k2;
    a.step = k2 / k1; a.kout = gcanet_knn_graph_columns(k1, k2);
    a.idx64 = idx64; a.idx32 = idx32; a.dist = nullptr; a.k_major = 0; a.index_base = 0;
    a.TR = pick_tile(C, N);
    if (metric == GCANET_METRIC_POINTS_NORMALS) {
        if (k2 > 168) return launch_bigk<METRIC_PN>(a, B, st);
        return launch_scan_k<6, METRIC_PN, 4>(a, B, st);
    }
    if (k2 > 168) return launch_bigk<METRIC_L2>(a, B, st);
    if (C == 3) return launch_scan_k<3, METRIC_L2, 4>(a, B, st);
    return launch_scan_k<0, METRIC_L2, 4>(a, B, st);
}

// knn_xyz.cu
bool knn_xyz_supported(int C, int N, int k2, int metric);
size_t knn_xyz_workspace_bytes(int B, int C, int N);
int knn_graph_xyz(const float *x, int B, int C, int N, int k1, int k2, int metric, int64_t *idx64, int32_t *idx32,
                  void *ws, float **norm_out, int **fallback_out, int unordered, cudaStream_t st);

// knn_tc.cu
size_t knn_tc_workspace_bytes(int B, int C, int N);
bool knn_tc_supported(int C, int N, int k2);
bool knn_tc_xyz_supported(int B, int N, int k2);
int knn_graph_tensor_cores(const float *x, int B, int C, int N, int k1, int k2, int64_t *idx64, int32_t *idx32,
                           void *ws, int unordered, int no_prune, cudaStream_t st);

}  // namespace gcanet

using namespace gcanet;

extern "C" int gcanet_knn_graph_columns(int k1, int k2) {
    if (k1 < 1 || k2 < k1) return 0;
    int step = k2 / k1;
    return (k2 + step - 1) / step;
}

static bool use_tensor_cores(int B, int C, int N, int k2, int metric) {
    const bool no_prune = (metric & GCANET_KNN_FLAG_NO_PRUNE) != 0;
    metric &= ~(GCANET_KNN_FLAG_UNORDERED | GCANET_KNN_FLAG_NO_PRUNE);
    if (metric != GCANET_METRIC_L2) return false;                      // the brute-force flag makes this false
    // xyz clouds: one MMA per key tile on the box-pruned scan; with GCANET_KNN_FLAG_NO_PRUNE (A/B tests) and outside that
    // scan's limits the CUDA-core kernel of knn_xyz.cu takes them
    if (C == 3) return !no_prune && knn_tc_xyz_supported(B, N, k2);
    return knn_tc_supported(C, N, k2);
}

extern "C" size_t gcanet_knn_graph_workspace_bytes(int B, int C, int N, int k2, int metric) {
    if (B < 1 || C < 1 || N < 1) return 0;
    if (use_tensor_cores(B, C, N, k2, metric)) return knn_tc_workspace_bytes(B, C, N);
    if (knn_xyz_supported(C, N, k2, metric & ~(GCANET_KNN_FLAG_UNORDERED | GCANET_KNN_FLAG_NO_PRUNE))) return knn_xyz_workspace_bytes(B, C, N);
    return align_up((size_t)B * N * sizeof(float));
}

extern "C" int gcanet_knn_graph(const float *x, int B, int C, int N, int k1, int k2, int metric,
                                int64_t *idx64, int32_t *idx32, void *ws, size_t ws_bytes,
                                gcanet_stream_t stream) {
    GCANET_REQUIRE(x != nullptr && (idx64 != nullptr || idx32 != nullptr), "knn_graph: null pointer");
    GCANET_REQUIRE(B >= 1 && C >= 1 && N >= 1, "knn_graph: bad shape B=%d C=%d N=%d", B, C, N);
    GCANET_REQUIRE(k1 >= 1 && k2 >= k1, "knn_graph: need 1 <= k1 <= k2 (k1=%d k2=%d)", k1, k2);
    GCANET_REQUIRE(k2 <= N, "knn_graph: k2=%d exceeds the number of points N=%d (topk would raise, M4:43)", k2, N);
    GCANET_REQUIRE(k2 <= 1024, "knn_graph: k2=%d > 1024 unsupported", k2);
    int flags = metric & ~0xff;
    GCANET_REQUIRE((flags & ~(GCANET_KNN_FLAG_BRUTE_FORCE | GCANET_KNN_FLAG_UNORDERED | GCANET_KNN_FLAG_NO_PRUNE)) == 0,
                   "knn_graph: unknown flag bits in metric 0x%x", metric);
    const int unordered = (flags & GCANET_KNN_FLAG_UNORDERED) ? 1 : 0;
    const int no_prune = (flags & GCANET_KNN_FLAG_NO_PRUNE) ? 1 : 0;
    flags &= ~(GCANET_KNN_FLAG_UNORDERED | GCANET_KNN_FLAG_NO_PRUNE);
    const int metric_in = metric;
    metric &= 0xff;
    GCANET_REQUIRE(metric == GCANET_METRIC_L2 || metric == GCANET_METRIC_POINTS_NORMALS, "knn_graph: bad metric %d", metric);
    GCANET_REQUIRE(metric != GCANET_METRIC_POINTS_NORMALS || C == 6,
                   "knn_graph: the points x normals metric needs C = 6 (got %d)", C);
    GCANET_REQUIRE(C <= 1024, "knn_graph: C=%d > 1024 unsupported", C);
    if (ws == nullptr || ws_bytes < gcanet_knn_graph_workspace_bytes(B, C, N, k2, metric_in) ||
        (reinterpret_cast<uintptr_t>(ws) % kAlign) != 0) {
        set_error("knn_graph: workspace too small or misaligned (%zu bytes given)", ws_bytes);
        return GCANET_ERR_WORKSPACE;
    }
    if (flags == 0 && use_tensor_cores(B, C, N, k2, metric | (no_prune ? GCANET_KNN_FLAG_NO_PRUNE : 0)))
        return knn_graph_tensor_cores(x, B, C, N, k1, k2, idx64, idx32, ws, unordered, no_prune, as_stream(stream));
    if (flags == 0 && knn_xyz_supported(C, N, k2, metric)) {
        float *norms = nullptr;
        int *cloud_fallback = nullptr;
        int rc = knn_graph_xyz(x, B, C, N, k1, k2, metric, idx64, idx32, ws, &norms, &cloud_fallback, unordered,
                               as_stream(stream));
        if (rc) return rc;
        // clouds the pruned scan declined (points x normals with non-unit normals): brute force, filtered per cloud
        return scan_self(x, norms, nullptr, nullptr, B, C, N, k1, k2, metric, idx64, idx32, as_stream(stream), cloud_fallback);
    }
    return knn_graph_cuda_cores(x, B, C, N, k1, k2, metric, idx64, idx32, static_cast<float *>(ws), as_stream(stream));
}

extern "C" size_t gcanet_knn_cuda_workspace_bytes(int batch, int dim, int ref_nb, int query_nb, int k) {
    (void)batch; (void)dim; (void)ref_nb; (void)query_nb; (void)k;
    return kAlign;   // the direct-difference metric needs no norms; kept for ABI symmetry
}

extern "C" int gcanet_knn_cuda(const float *ref, int ref_nb, const float *query, int query_nb, int dim, int k,
                               int batch, int index_base, float *dist, int64_t *ind, void *ws, size_t ws_bytes,
                               gcanet_stream_t stream) {
    (void)ws; (void)ws_bytes;
    GCANET_REQUIRE(ref && query && dist && ind, "knn_cuda: null pointer");
    GCANET_REQUIRE(batch >= 1 && dim >= 1 && ref_nb >= 1 && query_nb >= 1, "knn_cuda: bad shape");
    GCANET_REQUIRE(k >= 1 && k <= ref_nb, "knn_cuda: need 1 <= k <= ref_nb (k=%d ref_nb=%d)", k, ref_nb);
    GCANET_REQUIRE(k <= 1024 && dim <= 1024, "knn_cuda: k or dim > 1024 unsupported");
    GCANET_REQUIRE(index_base == 0 || index_base == 1, "knn_cuda: index_base must be 0 or 1");
    ScanArgs a{};
    a.ref = ref; a.qry = query; a.ref_norm = nullptr; a.qry_norm = nullptr;
    a.C = dim; a.Nr = ref_nb; a.Nq = query_nb; a.k = k; a.step = 1; a.kout = k;
    a.idx64 = ind; a.idx32 = nullptr; a.dist = dist; a.k_major = 1; a.index_base = index_base;
    a.TR = pick_tile(dim, ref_nb);
    cudaStream_t st = as_stream(stream);
    if (k > 168) return launch_bigk<METRIC_SSD>(a, batch, st);
    if (dim == 3) return launch_scan_k<3, METRIC_SSD, 4>(a, batch, st);
    return launch_scan_k<0, METRIC_SSD, 4>(a, batch, st);
}
