// Offset-prediction block of GCANet (OFFSET_PRED_MODULE + KPAM + cos_dist, M4:326-452), fused.
//
// Reference, per cloud: S = 120 key points (the same indices every call); cosine similarity of every point's instance
// feature to the keys' [N][S]; topk(k = 30) values + indices (computed twice, M4:421-422); gathers through N-fold repeats
// of the key tables ([B][N][120][128]); attention a = softmax_k(W2 relu(W1 d)) on the similarity values; edge tensor
// a_ik [f_j ; p_j - p_i] (131 channels) -> Conv2d(131 -> 128) -> GroupNorm(2) -> LeakyReLU -> max over k; concat with the
// point's own feature; Conv1d(256 -> 3).
//
// Here:  W [a (f_j ; p_j - p_i)] = a (T_j - q_i)  with  T_j = W_f f_j + W_p p_j  (120 x 128 per cloud, kept in shared
// memory) and q_i = W_p p_i, so one warp per point does similarity, ranking, attention and the 30 edges out of shared
// memory; GroupNorm + LeakyReLU commute with the max as in edgeconv.cu (sign of gamma picks max or min).  Nothing of
// size N x k x C or N x S x C is ever stored.  Backward recomputes the edges; the key-side gradients (dT, d k^) are
// vector reductions into global memory (resolved in L2).
#include "common.cuh"

#include <math_constants.h>

namespace gcanet {

constexpr unsigned OFULL = 0xffffffffu;
constexpr int OP_F = 128;            // feature channels = output channels of the edge conv
constexpr int OP_WARPS = 8;
constexpr int OP_PTS = 64;           // points per CTA (forward / reduce passes)
constexpr int OP_PTS_BWD = 128;      // points per CTA in the main backward pass (amortises the flush of the key gradients)
constexpr int OP_SMAX = 128;         // keys per cloud <= 128 (4 per lane)
constexpr int OP_KMAX = 32;          // neighbours per point <= 32 (one per lane)
constexpr int OP_KS = 4;             // padding of a key row in shared memory: rows of E + 4 floats stay 16-byte aligned and a quarter
                                     // warp reading four consecutive floats of eight different rows touches 32 different banks

struct OpArgs {
    const float *points;   // [B][N][3]
    const float *feat;     // [B][N][128]
    const float *inst;     // [B][N][E]
    const int *sub;        // [S]
    const float *cw;       // [128][131]
    const float *gamma, *beta;
    const float *w1, *w2;  // [k][k]
    const float *ow, *ob;  // [3][256], [3]
    float *T;              // [B][S][128]
    float *keyn;           // [B][S][E] normalised key instance features
    float *knorm;          // [B][S]
    unsigned char *selj;   // [B][N][32]
    float *seld, *sela;    // [B][N][32]
    float *ysel;           // [B][N][128]
    unsigned char *arg;    // [B][N][128]
    double *part;          // [B][nblk][G][2]
    float *stats;          // [B][G][2]
    float *out;            // [B][3][N]
    int B, N, S, k, E, G;
    float eps, slope;
};

// per cloud: T_j = W_f f_j + W_p p_j, normalised key instance features.   grid (S, B), block 128
__global__ void __launch_bounds__(128) op_keys_kernel(OpArgs a) {
    __shared__ float fj[OP_F + 3];
    __shared__ float red[4];
    const int j = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
    const size_t row = (size_t)b * a.N + a.sub[j];
    fj[t] = a.feat[row * OP_F + t];
    if (t < 3) fj[OP_F + t] = a.points[row * 3 + t];
    // norm of the key's instance feature
    float s = 0.f;
    for (int c = t; c < a.E; c += 128) { const float v = a.inst[row * a.E + c]; s = fmaf(v, v, s); }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(OFULL, s, o);
    if ((t & 31) == 0) red[t >> 5] = s;
    __syncthreads();
    const float nrm = sqrtf(red[0] + red[1] + red[2] + red[3]);
    for (int c = t; c < a.E; c += 128) a.keyn[((size_t)b * a.S + j) * a.E + c] = a.inst[row * a.E + c] / nrm;
    if (t == 0) a.knorm[(size_t)b * a.S + j] = nrm;
    const float *w = a.cw + (size_t)t * (OP_F + 3);
    float acc = 0.f;
#pragma unroll 8
    for (int f = 0; f < OP_F + 3; ++f) acc = fmaf(w[f], fj[f], acc);
    a.T[((size_t)b * a.S + j) * OP_F + t] = acc;
}

// shared-memory layout shared by the forward and backward point kernels
struct OpSmem {
    float *keyn;   // [S][E + OP_KS]
    float *T;      // [S][128]
    float *w1, *w2;  // [k][k + 1]
    float *wsc;    // per warp: sims[128] | u[E] | seld[32] | sela[32] | sh[32] | selj[32] (as int)
    int per_warp;
};

__device__ __forceinline__ OpSmem op_carve(float *base, int S, int E, int k) {
    OpSmem s;
    s.keyn = base;
    s.T = s.keyn + S * (E + OP_KS);
    s.w1 = s.T + S * OP_F;
    s.w2 = s.w1 + k * (k + 1);
    s.wsc = s.w2 + k * (k + 1);
    s.per_warp = OP_SMAX + E + 4 * 32;
    return s;
}
static size_t op_smem_floats(int S, int E, int k) { return (size_t)S * (E + OP_KS) + (size_t)S * OP_F + 2 * k * (k + 1) + OP_WARPS * (OP_SMAX + E + 4 * 32); }

__device__ __forceinline__ void op_load_tables(const OpArgs &a, const OpSmem &sm, int b) {
    for (int e = threadIdx.x; e < a.S * a.E; e += blockDim.x) sm.keyn[(e / a.E) * (a.E + OP_KS) + e % a.E] = a.keyn[(size_t)b * a.S * a.E + e];
    for (int e = threadIdx.x; e < a.S * OP_F; e += blockDim.x) sm.T[e] = a.T[(size_t)b * a.S * OP_F + e];
    for (int e = threadIdx.x; e < a.k * a.k; e += blockDim.x) {
        sm.w1[(e / a.k) * (a.k + 1) + e % a.k] = a.w1[e];
        sm.w2[(e / a.k) * (a.k + 1) + e % a.k] = a.w2[e];
    }
}

// similarity of point i to every key, ranking of the k most similar (descending, ties by key index), attention weights.
// Leaves: su = normalised instance feature, sd[r] = r-th similarity, sj[r] = its key, sa[r] = attention, sh[r] = hidden unit.
// Returns the norm of the point's instance feature.
__device__ __forceinline__ float op_select_attend(const OpArgs &a, const OpSmem &sm, float *ssim, float *su, float *sd, float *sa,
                                                  float *sh, int *sj, size_t row, int lane) {
    float nn = 0.f;
    for (int c = lane; c < a.E; c += 32) { const float v = a.inst[row * a.E + c]; su[c] = v; nn = fmaf(v, v, nn); }
    for (int o = 16; o; o >>= 1) nn += __shfl_xor_sync(OFULL, nn, o);
    const float nrm = sqrtf(nn);
    __syncwarp();
    for (int c = lane; c < a.E; c += 32) su[c] = su[c] / nrm;
    __syncwarp();
    float mine[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int key = lane + 32 * t;
        float v = -CUDART_INF_F;
        if (key < a.S) {
            const float4 *kr = reinterpret_cast<const float4 *>(sm.keyn + key * (a.E + OP_KS));
            const float4 *ur = reinterpret_cast<const float4 *>(su);
            float dot = 0.f;
            for (int c = 0; c < a.E / 4; ++c) {                 // same summation order as a scalar loop over the channels
                const float4 u4 = ur[c], k4 = kr[c];
                dot = fmaf(u4.x, k4.x, dot); dot = fmaf(u4.y, k4.y, dot); dot = fmaf(u4.z, k4.z, dot); dot = fmaf(u4.w, k4.w, dot);
            }
            v = -(1.f - dot);                                  // M4:340-341
        }
        mine[t] = v;
    }
    // The k most similar keys, descending, ties by key index.  Ranking every key against all S costs S x 4 compares per
    // lane; instead: the k-th largest similarity V by bisection over the order-preserving bit patterns (the warp counts
    // with one redux per probe and stops as soon as exactly k keys lie at or above the probe), the keys above V plus the
    // lowest-index ties are compacted into the k slots, and only those k are ranked against each other.
    unsigned key4[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const unsigned u = __float_as_uint(mine[t]);
        key4[t] = (lane + 32 * t < a.S) ? ((u & 0x80000000u) ? ~u : (u | 0x80000000u)) : 0u;   // absent keys sort below everything
    }
    unsigned klo, khi;                                           // count(key >= klo) >= k > count(key >= khi)
    {
        unsigned mn = 0xffffffffu, mx = 0u;
#pragma unroll
        for (int t = 0; t < 4; ++t) if (lane + 32 * t < a.S) { mn = min(mn, key4[t]); mx = max(mx, key4[t]); }
        klo = __reduce_min_sync(OFULL, mn);
        khi = __reduce_max_sync(OFULL, mx) + 1u;                 // (similarities are finite: no wrap-around)
    }
    int c_hi = 0;                                                // count(key >= khi)
    while (khi - klo > 1u) {
        const unsigned km = klo + ((khi - klo) >> 1);
        const int c = __reduce_add_sync(OFULL, (key4[0] >= km) + (key4[1] >= km) + (key4[2] >= km) + (key4[3] >= km));
        if (c >= a.k) { klo = km; if (c == a.k) break; } else { khi = km; c_hi = c; }
    }
    // klo: every key >= klo is selected when exactly k of them exist; otherwise klo is the k-th largest value itself
    // (khi = klo + 1), the c_hi keys above it are selected and the remaining k - c_hi come from the ties, lowest index first
    {
        const int cnt_ge = __reduce_add_sync(OFULL, (key4[0] >= klo) + (key4[1] >= klo) + (key4[2] >= klo) + (key4[3] >= klo));
        const bool exact = cnt_ge == a.k;
        int base = 0, ties_left = exact ? 0 : a.k - c_hi;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const bool valid = lane + 32 * t < a.S;
            const bool above = valid && (exact ? key4[t] >= klo : key4[t] > klo);
            const bool tie = valid && !exact && key4[t] == klo;
            const unsigned mt = __ballot_sync(OFULL, tie);
            const bool take_tie = tie && __popc(mt & ((1u << lane) - 1)) < ties_left;
            ties_left = max(0, ties_left - __popc(mt));
            const bool sel = above || take_tie;
            const unsigned ms = __ballot_sync(OFULL, sel);
            if (sel) {
                const int p = base + __popc(ms & ((1u << lane) - 1));
                ssim[p] = mine[t];                               // similarity scratch reused: the selected values ...
                sj[p] = lane + 32 * t;                           // ... and their keys, unordered
            }
            base += __popc(ms);
        }
    }
    __syncwarp();
    {
        float v = 0.f;
        int j = 0, rank = 0;
        if (lane < a.k) {
            v = ssim[lane]; j = sj[lane];
            for (int e = 0; e < a.k; ++e) {
                const float o = ssim[e];
                const int oj = sj[e];
                rank += (o > v || (o == v && oj < j)) ? 1 : 0;
            }
        }
        __syncwarp();
        if (lane < a.k) { sd[rank] = v; sj[rank] = j; }
    }
    __syncwarp();
    // attention (KPAM, M4:351-373): softmax over the k neighbours of W2 relu(W1 d)
    float z = 0.f;
    if (lane < a.k)
        for (int c = 0; c < a.k; ++c) z = fmaf(sm.w1[lane * (a.k + 1) + c], sd[c], z);
    sh[lane] = lane < a.k ? fmaxf(z, 0.f) : 0.f;
    __syncwarp();
    float pre = -CUDART_INF_F;
    if (lane < a.k) {
        pre = 0.f;
        for (int c = 0; c < a.k; ++c) pre = fmaf(sm.w2[lane * (a.k + 1) + c], sh[c], pre);
    }
    float mx = pre;
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(OFULL, mx, o));
    const float e = lane < a.k ? expf(pre - mx) : 0.f;
    float sum = e;
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(OFULL, sum, o);
    sa[lane] = e / sum;
    __syncwarp();
    return nrm;
}

// forward: one warp per point.  grid (ceil(N / OP_PTS), B), block 256, dynamic smem op_smem_floats
__global__ void __launch_bounds__(OP_WARPS * 32) op_forward_kernel(OpArgs a) {
    extern __shared__ __align__(16) float op_sm[];
    const OpSmem sm = op_carve(op_sm, a.S, a.E, a.k);
    const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    op_load_tables(a, sm, b);
    __syncthreads();
    float *ws = sm.wsc + warp * sm.per_warp;
    float *ssim = ws, *su = ws + OP_SMAX, *sd = su + a.E, *sa = sd + 32, *sh = sa + 32;
    int *sj = reinterpret_cast<int *>(sh + 32);
    const int c0 = lane * 4;
    float wp[4][3], sg[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
#pragma unroll
        for (int d = 0; d < 3; ++d) wp[v][d] = a.cw[(size_t)(c0 + v) * (OP_F + 3) + OP_F + d];
        sg[v] = a.gamma[c0 + v] < 0.f ? -1.f : 1.f;
    }
    double s1 = 0.0, s2 = 0.0;
    const int per_warp = OP_PTS / OP_WARPS;
    for (int pi = 0; pi < per_warp; ++pi) {
        const int i = blockIdx.x * OP_PTS + warp * per_warp + pi;
        if (i >= a.N) break;
        const size_t row = (size_t)b * a.N + i;
        op_select_attend(a, sm, ssim, su, sd, sa, sh, sj, row, lane);
        const float px = a.points[row * 3], py = a.points[row * 3 + 1], pz = a.points[row * 3 + 2];
        float q[4], zmax[4], vs[4], vq[4];
        int kb[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            q[v] = fmaf(wp[v][2], pz, fmaf(wp[v][1], py, wp[v][0] * px));
            zmax[v] = -CUDART_INF_F; vs[v] = 0.f; vq[v] = 0.f; kb[v] = 0;
        }
        for (int kk = 0; kk < a.k; ++kk) {
            const float at = sa[kk];
            const float4 t4 = *reinterpret_cast<const float4 *>(sm.T + sj[kk] * OP_F + c0);
            const float tv[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const float y = at * (tv[v] - q[v]);
                const float z = sg[v] * y;
                if (z > zmax[v]) { zmax[v] = z; kb[v] = kk; }
                vs[v] += y;
                vq[v] = fmaf(y, y, vq[v]);
            }
        }
        *reinterpret_cast<float4 *>(a.ysel + row * OP_F + c0) = make_float4(sg[0] * zmax[0], sg[1] * zmax[1], sg[2] * zmax[2], sg[3] * zmax[3]);
        *reinterpret_cast<uchar4 *>(a.arg + row * OP_F + c0) = make_uchar4(kb[0], kb[1], kb[2], kb[3]);
        a.selj[row * 32 + lane] = lane < a.k ? (unsigned char)sj[lane] : 0;
        a.seld[row * 32 + lane] = lane < a.k ? sd[lane] : 0.f;
        a.sela[row * 32 + lane] = lane < a.k ? sa[lane] : 0.f;
#pragma unroll
        for (int v = 0; v < 4; ++v) { s1 += (double)vs[v]; s2 += (double)vq[v]; }
        __syncwarp();
    }
    // per-group sums: the lanes of a group are contiguous (channels of a lane never straddle a group)
    const int lpg = 32 / a.G;
    for (int o = 1; o < lpg; o <<= 1) { s1 += __shfl_xor_sync(OFULL, s1, o); s2 += __shfl_xor_sync(OFULL, s2, o); }
    __syncwarp();
    double *wred = reinterpret_cast<double *>(ws);               // the warp's similarity scratch (128 floats) is free now
    if (lane % lpg == 0) { wred[(lane / lpg) * 2] = s1; wred[(lane / lpg) * 2 + 1] = s2; }
    __syncthreads();
    if (threadIdx.x < a.G * 2) {
        const int g = threadIdx.x >> 1, which = threadIdx.x & 1;
        double s = 0.0;
        for (int w = 0; w < OP_WARPS; ++w) s += reinterpret_cast<const double *>(sm.wsc + w * sm.per_warp)[g * 2 + which];
        a.part[(((size_t)b * gridDim.x + blockIdx.x) * a.G + g) * 2 + which] = s;
    }
}

__global__ void op_stats_kernel(const double *__restrict__ part, float *__restrict__ stats, int nblk, int G, double count, float eps) {
    const int b = blockIdx.x, g = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (g >= G) return;
    double s1 = 0.0, s2 = 0.0;
    for (int i = lane; i < nblk; i += 32) {
        s1 += part[(((size_t)b * nblk + i) * G + g) * 2];
        s2 += part[(((size_t)b * nblk + i) * G + g) * 2 + 1];
    }
    for (int o = 16; o; o >>= 1) { s1 += __shfl_xor_sync(OFULL, s1, o); s2 += __shfl_xor_sync(OFULL, s2, o); }
    if (lane == 0) {
        const double mean = s1 / count;
        double var = s2 / count - mean * mean;
        if (var < 0.0) var = 0.0;
        stats[((size_t)b * G + g) * 2] = (float)mean;
        stats[((size_t)b * G + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
    }
}

// feat_att = LReLU(GN(ysel)); offsets = W_o [feat_att ; feature] + b_o      one warp per point
__global__ void __launch_bounds__(256) op_finish_kernel(OpArgs a) {
    const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + warp;
    if (i >= a.N) return;
    const int c0 = lane * 4, g = c0 / (OP_F / a.G);
    const float mean = a.stats[((size_t)b * a.G + g) * 2], rstd = a.stats[((size_t)b * a.G + g) * 2 + 1];
    const size_t row = (size_t)b * a.N + i;
    const float4 ys = *reinterpret_cast<const float4 *>(a.ysel + row * OP_F + c0);
    const float4 ft = *reinterpret_cast<const float4 *>(a.feat + row * OP_F + c0);
    const float yv[4] = {ys.x, ys.y, ys.z, ys.w}, fv[4] = {ft.x, ft.y, ft.z, ft.w};
    float o[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        const float u = (yv[v] - mean) * rstd * a.gamma[c0 + v] + a.beta[c0 + v];
        const float fa = u > 0.f ? u : u * a.slope;
#pragma unroll
        for (int d = 0; d < 3; ++d) o[d] = fmaf(a.ow[d * 2 * OP_F + OP_F + c0 + v], fv[v], fmaf(a.ow[d * 2 * OP_F + c0 + v], fa, o[d]));
    }
#pragma unroll
    for (int d = 0; d < 3; ++d)
        for (int off = 16; off; off >>= 1) o[d] += __shfl_xor_sync(OFULL, o[d], off);
    if (lane < 3) a.out[((size_t)b * 3 + lane) * a.N + i] = (lane == 0 ? o[0] : (lane == 1 ? o[1] : o[2])) + a.ob[lane];
}

// ------------------------------------------------------------------------------------------------ backward
struct OpBwdArgs {
    OpArgs f;
    const float *gout;     // [B][3][N]
    float *du;             // [B][N][128]   d(loss)/d(pre-activation of LeakyReLU)
    float *rpart;          // [B][nblk][128][2]  per-CTA sums of du, du * yhat
    float *owpart;         // [B * nblk][3 * 256 + 3] per-CTA partials of d(W_o), d(b_o)
    float *coef;           // [B][G][2]
    float *dfeat;          // [B][N][128]
    float *dinst;          // [B][N][E]
    float *dT;             // [B][S][128]   zeroed
    float *dkn;            // [B][S][E]     zeroed
    float *dw1, *dw2;      // [k][k]        zeroed
    float *dwp;            // [128][3]      zeroed: gradient of the position columns of the conv through q_i
};

// pass 1: direct feature gradient, du, GroupNorm sums, W_o / b_o partials.   one warp per point, 8 points per warp
__global__ void __launch_bounds__(OP_WARPS * 32) op_bwd_reduce_kernel(OpBwdArgs p) {
    const OpArgs &a = p.f;
    __shared__ float red[OP_WARPS][OP_F][2];
    __shared__ float redw[OP_WARPS][3 * 2 * OP_F + 3];
    const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c0 = lane * 4, g = c0 / (OP_F / a.G);
    const float mean = a.stats[((size_t)b * a.G + g) * 2], rstd = a.stats[((size_t)b * a.G + g) * 2 + 1];
    float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0}, dwo[3][8], dbo[3] = {0, 0, 0};
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int v = 0; v < 8; ++v) dwo[d][v] = 0.f;
    const int per_warp = OP_PTS / OP_WARPS;
    for (int pi = 0; pi < per_warp; ++pi) {
        const int i = blockIdx.x * OP_PTS + warp * per_warp + pi;
        if (i >= a.N) break;
        const size_t row = (size_t)b * a.N + i;
        float gd[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) gd[d] = p.gout[((size_t)b * 3 + d) * a.N + i];
        const float4 ys = *reinterpret_cast<const float4 *>(a.ysel + row * OP_F + c0);
        const float4 ft = *reinterpret_cast<const float4 *>(a.feat + row * OP_F + c0);
        const float yv[4] = {ys.x, ys.y, ys.z, ys.w}, fv[4] = {ft.x, ft.y, ft.z, ft.w};
        float duv[4], dfv[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const float yh = (yv[v] - mean) * rstd;
            const float u = yh * a.gamma[c0 + v] + a.beta[c0 + v];
            const float fa = u > 0.f ? u : u * a.slope;
            float dfa = 0.f, df = 0.f;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                dfa = fmaf(a.ow[d * 2 * OP_F + c0 + v], gd[d], dfa);
                df = fmaf(a.ow[d * 2 * OP_F + OP_F + c0 + v], gd[d], df);
                dwo[d][v] = fmaf(gd[d], fa, dwo[d][v]);
                dwo[d][4 + v] = fmaf(gd[d], fv[v], dwo[d][4 + v]);
            }
            duv[v] = u > 0.f ? dfa : dfa * a.slope;
            dfv[v] = df;
            s1[v] += duv[v];
            s2[v] = fmaf(duv[v], yh, s2[v]);
        }
        if (lane == 0) { dbo[0] += gd[0]; dbo[1] += gd[1]; dbo[2] += gd[2]; }
        *reinterpret_cast<float4 *>(p.du + row * OP_F + c0) = make_float4(duv[0], duv[1], duv[2], duv[3]);
        *reinterpret_cast<float4 *>(p.dfeat + row * OP_F + c0) = make_float4(dfv[0], dfv[1], dfv[2], dfv[3]);
    }
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        red[warp][c0 + v][0] = s1[v]; red[warp][c0 + v][1] = s2[v];
#pragma unroll
        for (int d = 0; d < 3; ++d) { redw[warp][d * 2 * OP_F + c0 + v] = dwo[d][v]; redw[warp][d * 2 * OP_F + OP_F + c0 + v] = dwo[d][4 + v]; }
    }
    if (lane == 0) { redw[warp][6 * OP_F] = dbo[0]; redw[warp][6 * OP_F + 1] = dbo[1]; redw[warp][6 * OP_F + 2] = dbo[2]; }
    __syncthreads();
    const size_t blk = (size_t)b * gridDim.x + blockIdx.x;
    for (int e = threadIdx.x; e < OP_F * 2; e += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < OP_WARPS; ++w) s += red[w][e >> 1][e & 1];
        p.rpart[blk * OP_F * 2 + e] = s;
    }
    for (int e = threadIdx.x; e < 6 * OP_F + 3; e += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < OP_WARPS; ++w) s += redw[w][e];
        p.owpart[blk * (6 * OP_F + 3) + e] = s;
    }
}

// per cloud: (A_g, K_g) and the per-channel sums behind dgamma, dbeta.   grid B CTAs of 256 threads
__global__ void __launch_bounds__(256) op_bwd_coef_kernel(OpBwdArgs p, float *__restrict__ dgamma, float *__restrict__ dbeta,
                                                          float *__restrict__ dow, float *__restrict__ dob, double *__restrict__ sbc,
                                                          int nblk) {
    const OpArgs &a = p.f;
    __shared__ double gs[2][32];
    const int b = blockIdx.x;
    for (int e = threadIdx.x; e < 2 * 32; e += blockDim.x) gs[e / 32][e % 32] = 0.0;
    __syncthreads();
    for (int c = threadIdx.x; c < OP_F; c += blockDim.x) {
        double s1 = 0.0, s2 = 0.0;
        for (int i = 0; i < nblk; ++i) {
            s1 += (double)p.rpart[((size_t)b * nblk + i) * OP_F * 2 + c * 2];
            s2 += (double)p.rpart[((size_t)b * nblk + i) * OP_F * 2 + c * 2 + 1];
        }
        sbc[((size_t)b * OP_F + c) * 2] = s1;
        sbc[((size_t)b * OP_F + c) * 2 + 1] = s2;
        const int g = c / (OP_F / a.G);
        atomicAdd(&gs[0][g], s1 * (double)a.gamma[c]);
        atomicAdd(&gs[1][g], s2 * (double)a.gamma[c]);
    }
    __syncthreads();
    if (threadIdx.x < a.G) {
        const int g = threadIdx.x;
        const double cnt = (double)(OP_F / a.G) * a.N * a.k;
        const double m1 = gs[0][g] / cnt, m2 = gs[1][g] / cnt;
        const double mean = a.stats[((size_t)b * a.G + g) * 2], rstd = a.stats[((size_t)b * a.G + g) * 2 + 1];
        p.coef[((size_t)b * a.G + g) * 2] = (float)(-rstd * m1 + rstd * rstd * m2 * mean);
        p.coef[((size_t)b * a.G + g) * 2 + 1] = (float)(-rstd * rstd * m2);
    }
    (void)dgamma; (void)dbeta; (void)dow; (void)dob;
}

// d(W_o), d(b_o): sum of the per-CTA partials in a fixed order.  block (32 entries, 32 slices of the partial list)
__global__ void __launch_bounds__(1024) op_bwd_ow_kernel(const float *__restrict__ owpart, int total, float *__restrict__ dow,
                                                         float *__restrict__ dob) {
    __shared__ double red[32][33];
    const int e = blockIdx.x * 32 + threadIdx.x;
    double s = 0.0;
    if (e < 6 * OP_F + 3)
        for (int i = threadIdx.y; i < total; i += 32) s += (double)owpart[(size_t)i * (6 * OP_F + 3) + e];
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && e < 6 * OP_F + 3) {
        double t = 0.0;
        for (int r = 0; r < 32; ++r) t += red[r][threadIdx.x];
        if (e < 6 * OP_F) dow[e] = (float)t; else dob[e - 6 * OP_F] = (float)t;
    }
}

__global__ void op_bwd_affine_kernel(const double *__restrict__ sbc, float *__restrict__ dgamma, float *__restrict__ dbeta, int B) {
    const int c = threadIdx.x;
    double s1 = 0.0, s2 = 0.0;
    for (int b = 0; b < B; ++b) { s1 += sbc[((size_t)b * OP_F + c) * 2]; s2 += sbc[((size_t)b * OP_F + c) * 2 + 1]; }
    dbeta[c] = (float)s1;
    dgamma[c] = (float)s2;
}

// pass 2: the edges again.  One warp per point.  Key-side gradients (dT, d k^) go straight to global memory as vector
// reductions (red.global.add.v4.f32 / v2.f32, resolved in L2): fp32 atomicAdd on SHARED memory is a compare-and-swap loop,
// and with 120 keys shared by all points it made this kernel 16 % of the whole training step (959 M instructions, ncu).
__global__ void __launch_bounds__(OP_WARPS * 32, 2) op_bwd_main_kernel(OpBwdArgs p) {
    const OpArgs &a = p.f;
    extern __shared__ __align__(16) float op_sm[];
    __shared__ float s_red[OP_WARPS][2 * OP_KMAX + 1];
    const OpSmem sm = op_carve(op_sm, a.S, a.E, a.k);
    const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    op_load_tables(a, sm, b);
    __syncthreads();
    float *ws = sm.wsc + warp * sm.per_warp;
    float *ssim = ws, *su = ws + OP_SMAX, *sd = su + a.E, *sa = sd + 32, *sh = sa + 32;
    int *sj = reinterpret_cast<int *>(sh + 32);
    const int c0 = lane * 4, g = c0 / (OP_F / a.G);
    const float rstd = a.stats[((size_t)b * a.G + g) * 2 + 1];
    const float Ag = p.coef[((size_t)b * a.G + g) * 2], Kg = p.coef[((size_t)b * a.G + g) * 2 + 1];
    float wp[4][3], gm[4], dwp[4][3];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        gm[v] = a.gamma[c0 + v];
#pragma unroll
        for (int d = 0; d < 3; ++d) { wp[v][d] = a.cw[(size_t)(c0 + v) * (OP_F + 3) + OP_F + d]; dwp[v][d] = 0.f; }
    }
    // attention-weight gradients: lane r keeps row r of dW2 (dpre_r h_c) and row r of dW1 (dz_r d_c) in registers
    float dw1r[OP_KMAX], dw2r[OP_KMAX];
#pragma unroll
    for (int c = 0; c < OP_KMAX; ++c) { dw1r[c] = 0.f; dw2r[c] = 0.f; }
    float *dT = p.dT + (size_t)b * a.S * OP_F + c0;
    float *dkn = p.dkn + (size_t)b * a.S * a.E;
    const int per_warp = OP_PTS_BWD / OP_WARPS;
    for (int pi = 0; pi < per_warp; ++pi) {
        const int i = blockIdx.x * OP_PTS_BWD + warp * per_warp + pi;
        if (i >= a.N) break;
        const size_t row = (size_t)b * a.N + i;
        // saved selection and attention (identical to what the forward used)
        sj[lane] = a.selj[row * 32 + lane];
        sd[lane] = a.seld[row * 32 + lane];
        sa[lane] = a.sela[row * 32 + lane];
        float nn = 0.f;
        for (int c = lane; c < a.E; c += 32) { const float v = a.inst[row * a.E + c]; su[c] = v; nn = fmaf(v, v, nn); }
        for (int o = 16; o; o >>= 1) nn += __shfl_xor_sync(OFULL, nn, o);
        const float nrm = sqrtf(nn);
        __syncwarp();
        for (int c = lane; c < a.E; c += 32) su[c] = su[c] / nrm;
        // hidden units of the attention MLP
        float z = 0.f;
        if (lane < a.k)
            for (int c = 0; c < a.k; ++c) z = fmaf(sm.w1[lane * (a.k + 1) + c], sd[c], z);
        sh[lane] = lane < a.k ? fmaxf(z, 0.f) : 0.f;
        __syncwarp();

        const float px = a.points[row * 3], py = a.points[row * 3 + 1], pz = a.points[row * 3 + 2];
        const float4 du4 = *reinterpret_cast<const float4 *>(p.du + row * OP_F + c0);
        const uchar4 ar4 = *reinterpret_cast<const uchar4 *>(a.arg + row * OP_F + c0);
        const float duv[4] = {du4.x, du4.y, du4.z, du4.w};
        const int ak[4] = {ar4.x, ar4.y, ar4.z, ar4.w};
        float q[4], sv[4], dq[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            q[v] = fmaf(wp[v][2], pz, fmaf(wp[v][1], py, wp[v][0] * px));
            sv[v] = rstd * gm[v] * duv[v];
            dq[v] = 0.f;
        }
        // d(attention): da_k = sum over the 128 channels of dy t.  One butterfly reduction per edge is a chain of five
        // dependent shuffles (150 per point); eight edges at a time are reduce-scattered instead: 7 + 2 shuffles per
        // chunk, after which lane l holds the total of edge 8 chunk + (l & 7)
        float da_mine = 0.f;
#pragma unroll
        for (int ch = 0; ch < OP_KMAX / 8; ++ch) {
            if (ch * 8 >= a.k) break;
            float dv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int kk = ch * 8 + u;
                dv[u] = 0.f;
                if (kk < a.k) {
                    const float at = sa[kk];
                    const int j = sj[kk];
                    const float4 t4 = *reinterpret_cast<const float4 *>(sm.T + j * OP_F + c0);
                    const float tv[4] = {t4.x, t4.y, t4.z, t4.w};
                    float dap = 0.f, adt[4];
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const float t = tv[v] - q[v];
                        const float dy = (ak[v] == kk ? sv[v] : 0.f) + fmaf(Kg, at * t, Ag);
                        dap = fmaf(dy, t, dap);
                        adt[v] = at * dy;
                        dq[v] -= adt[v];
                    }
                    atomicAdd(reinterpret_cast<float4 *>(dT + (size_t)j * OP_F), make_float4(adt[0], adt[1], adt[2], adt[3]));
                    dv[u] = dap;
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float send = (lane & 4) ? dv[i] : dv[i + 4], keep = (lane & 4) ? dv[i + 4] : dv[i];
                dv[i] = keep + __shfl_xor_sync(OFULL, send, 4);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float send = (lane & 2) ? dv[i] : dv[i + 2], keep = (lane & 2) ? dv[i + 2] : dv[i];
                dv[i] = keep + __shfl_xor_sync(OFULL, send, 2);
            }
            {
                const float send = (lane & 1) ? dv[0] : dv[1], keep = (lane & 1) ? dv[1] : dv[0];
                dv[0] = keep + __shfl_xor_sync(OFULL, send, 1);
            }
            dv[0] += __shfl_xor_sync(OFULL, dv[0], 8);
            dv[0] += __shfl_xor_sync(OFULL, dv[0], 16);
            if ((lane >> 3) == ch) da_mine = dv[0];
        }
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            dwp[v][0] = fmaf(dq[v], px, dwp[v][0]);
            dwp[v][1] = fmaf(dq[v], py, dwp[v][1]);
            dwp[v][2] = fmaf(dq[v], pz, dwp[v][2]);
        }
        // softmax backward
        const float at = lane < a.k ? sa[lane] : 0.f;
        float dot = at * da_mine;
        for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(OFULL, dot, o);
        const float dpre = at * (da_mine - dot);
        // MLP backward: dh = W2^T dpre, dz = dh [z > 0], dd = W1^T dz
        ssim[lane] = dpre;                                   // similarity scratch reused: [0, 32) dpre, [32, 64) dz, [64, 96) dd
        __syncwarp();
        float dh = 0.f;
        if (lane < a.k)
            for (int r = 0; r < a.k; ++r) dh = fmaf(sm.w2[r * (a.k + 1) + lane], ssim[r], dh);
        const float dz = (lane < a.k && z > 0.f) ? dh : 0.f;
        ssim[32 + lane] = dz;
        __syncwarp();
        float dd = 0.f;
        if (lane < a.k)
            for (int c = 0; c < a.k; ++c) dd = fmaf(sm.w1[c * (a.k + 1) + lane], ssim[32 + c], dd);
#pragma unroll
        for (int c = 0; c < OP_KMAX; ++c) {                  // (lanes >= k carry dpre = dz = 0, columns >= k carry sh = sd = 0)
            dw2r[c] = fmaf(dpre, sh[c], dw2r[c]);
            dw1r[c] = fmaf(dz, sd[c], dw1r[c]);
        }
        ssim[64 + lane] = dd;                                // gradient w.r.t. the similarity of the lane-th neighbour
        __syncwarp();
        // cosine similarity backward: sim_j = u^ . k^_j - 1.  A lane owns the component pairs (64 e + 2 lane, + 1) of the
        // instance feature (E <= 256: at most 4 pairs); d u^ = sum_k dd_k k^_{j_k}; every key collects dd_k u^ (one
        // two-float reduction per lane, key and pair)
        float dun[4][2];
        float dotu = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int c = 64 * e + 2 * lane;
            dun[e][0] = 0.f; dun[e][1] = 0.f;
            if (c < a.E) {
                const float u0 = su[c], u1 = su[c + 1];
                for (int kk = 0; kk < a.k; ++kk) {
                    const float ddk = ssim[64 + kk];
                    const float *kr = sm.keyn + sj[kk] * (a.E + OP_KS) + c;
                    dun[e][0] = fmaf(ddk, kr[0], dun[e][0]);
                    dun[e][1] = fmaf(ddk, kr[1], dun[e][1]);
                    atomicAdd(reinterpret_cast<float2 *>(dkn + (size_t)sj[kk] * a.E + c), make_float2(ddk * u0, ddk * u1));
                }
                dotu = fmaf(dun[e][0], u0, fmaf(dun[e][1], u1, dotu));
            }
        }
        for (int o = 16; o; o >>= 1) dotu += __shfl_xor_sync(OFULL, dotu, o);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int c = 64 * e + 2 * lane;
            if (c < a.E)                                     // (I - u^ u^T) d u^ / |u|
                *reinterpret_cast<float2 *>(p.dinst + row * a.E + c) =
                    make_float2((dun[e][0] - dotu * su[c]) / nrm, (dun[e][1] - dotu * su[c + 1]) / nrm);
        }
        __syncwarp();
    }
    // per-thread accumulators -> global (one reduction per thread and entry, after a cross-warp sum for the k x k matrices)
#pragma unroll
    for (int v = 0; v < 4; ++v)
#pragma unroll
        for (int d = 0; d < 3; ++d) atomicAdd(p.dwp + (c0 + v) * 3 + d, dwp[v][d]);
#pragma unroll 1
    for (int c = 0; c < a.k; ++c) {
        float v1 = 0.f, v2 = 0.f;
#pragma unroll
        for (int cc = 0; cc < OP_KMAX; ++cc) { if (cc == c) { v1 = dw1r[cc]; v2 = dw2r[cc]; } }
        __syncthreads();
        if (lane < a.k) { s_red[warp][lane] = v1; s_red[warp][OP_KMAX + lane] = v2; }
        __syncthreads();
        if (warp == 0 && lane < a.k) {
            float t1 = 0.f, t2 = 0.f;
            for (int w = 0; w < OP_WARPS; ++w) { t1 += s_red[w][lane]; t2 += s_red[w][OP_KMAX + lane]; }
            atomicAdd(p.dw1 + lane * a.k + c, t1);          // dW1[row = lane][col = c]
            atomicAdd(p.dw2 + lane * a.k + c, t2);
        }
    }
}

// key-side epilogue: gradients that reached the key tables go back to the key points' rows.   grid (S, B), block 128
__global__ void __launch_bounds__(128) op_bwd_keys_kernel(OpBwdArgs p) {
    const OpArgs &a = p.f;
    __shared__ float dt[OP_F];
    __shared__ float red[4];
    const int j = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
    const size_t row = (size_t)b * a.N + a.sub[j];
    dt[t] = p.dT[((size_t)b * a.S + j) * OP_F + t];
    __syncthreads();
    // d feature[key] += W_f^T dT_j
    float acc = 0.f;
    for (int c = 0; c < OP_F; ++c) acc = fmaf(a.cw[(size_t)c * (OP_F + 3) + t], dt[c], acc);
    p.dfeat[row * OP_F + t] += acc;
    // d inst[key] += (I - k^ k^T) d k^ / |v|
    const float *kn = a.keyn + ((size_t)b * a.S + j) * a.E;
    const float *dk = p.dkn + ((size_t)b * a.S + j) * a.E;
    float s = 0.f;
    for (int c = t; c < a.E; c += 128) s = fmaf(kn[c], dk[c], s);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(OFULL, s, o);
    if ((t & 31) == 0) red[t >> 5] = s;
    __syncthreads();
    const float dot = red[0] + red[1] + red[2] + red[3];
    const float nrm = a.knorm[(size_t)b * a.S + j];
    for (int c = t; c < a.E; c += 128) p.dinst[row * a.E + c] += (dk[c] - dot * kn[c]) / nrm;
}

// d conv weight [128][131]: columns f < 128 and the p_j part of the position columns from dT, the -p_i part from dwp.
// dcw[c][f] = sum over (cloud, key) of dT[c] * [feature ; position][f]: grid (16 channel blocks, B), thread = column f;
// the per-cloud partials are summed in a fixed order by op_bwd_convw_sum_kernel.
__global__ void __launch_bounds__(OP_F + 3) op_bwd_convw_kernel(OpBwdArgs p, float *__restrict__ part) {
    const OpArgs &a = p.f;
    __shared__ int rows[OP_SMAX];
    __shared__ float dts[OP_SMAX][8];
    const int cb = blockIdx.x, b = blockIdx.y, f = threadIdx.x;
    for (int j = f; j < a.S; j += OP_F + 3) rows[j] = a.sub[j];
    for (int e = f; e < a.S * 8; e += OP_F + 3) dts[e >> 3][e & 7] = p.dT[((size_t)b * a.S + (e >> 3)) * OP_F + cb * 8 + (e & 7)];
    __syncthreads();
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int j = 0; j < a.S; ++j) {
        const size_t row = (size_t)b * a.N + rows[j];
        const float x = f < OP_F ? a.feat[row * OP_F + f] : a.points[row * 3 + f - OP_F];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = fmaf(dts[j][u], x, acc[u]);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) part[((size_t)b * OP_F + cb * 8 + u) * (OP_F + 3) + f] = acc[u];
}

__global__ void __launch_bounds__(OP_F + 3) op_bwd_convw_sum_kernel(OpBwdArgs p, const float *__restrict__ part, float *__restrict__ dcw) {
    const int c = blockIdx.x, f = threadIdx.x;
    double acc = 0.0;
    for (int b = 0; b < p.f.B; ++b) acc += (double)part[((size_t)b * OP_F + c) * (OP_F + 3) + f];
    if (f >= OP_F) acc += (double)p.dwp[c * 3 + f - OP_F];
    dcw[(size_t)c * (OP_F + 3) + f] = (float)acc;
}

// ------------------------------------------------------------------------------------------------ host side
struct OpSaved { float *T, *keyn, *knorm, *seld, *sela, *ysel, *stats; unsigned char *selj, *arg; };
struct OpWs { double *part, *sbc; float *du, *rpart, *owpart, *coef, *cwpart, *dT, *dkn, *dw, *dwp; };

static int op_check(const gcanet_offset_desc *d) {
    GCANET_REQUIRE(d != nullptr, "offset_pred: null descriptor");
    GCANET_REQUIRE(d->B >= 1 && d->B <= 65535 && d->N >= 1, "offset_pred: bad shape B=%d N=%d", d->B, d->N);
    GCANET_REQUIRE(d->S >= 4 && d->S <= OP_SMAX && d->S <= d->N && d->S % 4 == 0,
                   "offset_pred: S=%d key points must be a multiple of 4 in [4, min(128, N)]", d->S);
    GCANET_REQUIRE(d->k >= 1 && d->k <= OP_KMAX && d->k <= d->S, "offset_pred: k=%d must be in [1, min(32, S)]", d->k);
    GCANET_REQUIRE(d->E >= 4 && d->E <= 256 && d->E % 4 == 0, "offset_pred: E=%d instance channels must be a multiple of 4 in [4, 256]", d->E);
    GCANET_REQUIRE(d->groups >= 1 && OP_F % d->groups == 0 && 32 % d->groups == 0, "offset_pred: groups=%d must divide 32", d->groups);
    GCANET_REQUIRE(d->eps > 0.f && d->slope >= 0.f, "offset_pred: eps must be positive and negative_slope >= 0");
    GCANET_REQUIRE((long long)d->B * d->N < 2147483647ll, "offset_pred: B * N does not fit 32 bits");
    return GCANET_OK;
}

static size_t op_plan_saved(const gcanet_offset_desc *d, void *base, OpSaved *s) {
    Carver cv(base);
    const size_t bn = (size_t)d->B * d->N, bs = (size_t)d->B * d->S;
    float *T = cv.take<float>(bs * OP_F), *keyn = cv.take<float>(bs * d->E), *knorm = cv.take<float>(bs);
    float *seld = cv.take<float>(bn * 32), *sela = cv.take<float>(bn * 32), *ysel = cv.take<float>(bn * OP_F);
    float *stats = cv.take<float>((size_t)d->B * d->groups * 2);
    unsigned char *selj = cv.take<unsigned char>(bn * 32), *arg = cv.take<unsigned char>(bn * OP_F);
    if (s) { s->T = T; s->keyn = keyn; s->knorm = knorm; s->seld = seld; s->sela = sela; s->ysel = ysel; s->stats = stats; s->selj = selj; s->arg = arg; }
    return cv.off;
}

static size_t op_plan_ws(const gcanet_offset_desc *d, void *base, OpWs *w) {
    Carver cv(base);
    const size_t bn = (size_t)d->B * d->N, bs = (size_t)d->B * d->S;
    const int nblk = ceil_div(d->N, OP_PTS);
    double *part = cv.take<double>((size_t)d->B * nblk * d->groups * 2);
    double *sbc = cv.take<double>((size_t)d->B * OP_F * 2);
    float *du = cv.take<float>(bn * OP_F);
    float *rpart = cv.take<float>((size_t)d->B * nblk * OP_F * 2);
    float *owpart = cv.take<float>((size_t)d->B * nblk * (6 * OP_F + 3));
    float *coef = cv.take<float>((size_t)d->B * d->groups * 2);
    float *cwpart = cv.take<float>((size_t)d->B * OP_F * (OP_F + 3));
    // zeroed in one memset: dT | dkn | dw1 | dw2 | dwp
    float *dT = cv.take<float>(bs * OP_F + bs * d->E + 2 * d->k * d->k + OP_F * 3);
    if (w) { w->part = part; w->sbc = sbc; w->du = du; w->rpart = rpart; w->owpart = owpart; w->coef = coef; w->cwpart = cwpart; w->dT = dT;
             w->dkn = dT + bs * OP_F; w->dw = w->dkn + bs * d->E; w->dwp = w->dw + 2 * d->k * d->k; }
    return cv.off;
}

static OpArgs op_args(const gcanet_offset_desc *d, const float *points, const float *feature, const float *inst, const int *sub,
                      const float *cw, const float *gamma, const float *beta, const float *w1, const float *w2, const float *ow,
                      const float *ob, const OpSaved &sv, double *part, float *out) {
    OpArgs a{};
    a.points = points; a.feat = feature; a.inst = inst; a.sub = sub; a.cw = cw; a.gamma = gamma; a.beta = beta; a.w1 = w1; a.w2 = w2;
    a.ow = ow; a.ob = ob; a.T = sv.T; a.keyn = sv.keyn; a.knorm = sv.knorm; a.selj = sv.selj; a.seld = sv.seld; a.sela = sv.sela;
    a.ysel = sv.ysel; a.arg = sv.arg; a.part = part; a.stats = sv.stats; a.out = out;
    a.B = d->B; a.N = d->N; a.S = d->S; a.k = d->k; a.E = d->E; a.G = d->groups; a.eps = d->eps; a.slope = d->slope;
    return a;
}

}  // namespace gcanet

using namespace gcanet;

extern "C" size_t gcanet_offset_pred_saved_bytes(const gcanet_offset_desc *d) {
    if (op_check(d) != GCANET_OK) return 0;
    return op_plan_saved(d, nullptr, nullptr);
}

extern "C" size_t gcanet_offset_pred_workspace_bytes(const gcanet_offset_desc *d) {
    if (op_check(d) != GCANET_OK) return 0;
    return op_plan_ws(d, nullptr, nullptr);
}

extern "C" int gcanet_offset_pred_forward(const gcanet_offset_desc *d, const float *points, const float *feature, const float *inst,
                                          const int32_t *key_index, const float *conv_w, const float *gamma, const float *beta,
                                          const float *att_w1, const float *att_w2, const float *off_w, const float *off_b,
                                          float *out, void *saved, void *ws, size_t ws_bytes, gcanet_stream_t stream) {
    int rc = op_check(d);
    if (rc) return rc;
    GCANET_REQUIRE(points && feature && inst && key_index && conv_w && gamma && beta && att_w1 && att_w2 && off_w && off_b && out && saved,
                   "offset_pred_forward: null pointer");
    const size_t need = op_plan_ws(d, nullptr, nullptr);
    if (ws == nullptr || ws_bytes < need || reinterpret_cast<uintptr_t>(ws) % kAlign || reinterpret_cast<uintptr_t>(saved) % kAlign) {
        set_error("offset_pred_forward: workspace too small or misaligned (%zu given, %zu needed)", ws_bytes, need);
        return GCANET_ERR_WORKSPACE;
    }
    OpSaved sv; OpWs w;
    op_plan_saved(d, saved, &sv);
    op_plan_ws(d, ws, &w);
    cudaStream_t st = as_stream(stream);
    OpArgs a = op_args(d, points, feature, inst, key_index, conv_w, gamma, beta, att_w1, att_w2, off_w, off_b, sv, w.part, out);
    op_keys_kernel<<<dim3(d->S, d->B), 128, 0, st>>>(a);
    GCANET_LAUNCH_OK("op_keys_kernel");
    const size_t smem = op_smem_floats(d->S, d->E, d->k) * sizeof(float);
    GCANET_CUDA_OK(cudaFuncSetAttribute(op_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int nblk = ceil_div(d->N, OP_PTS);
    op_forward_kernel<<<dim3(nblk, d->B), OP_WARPS * 32, smem, st>>>(a);
    GCANET_LAUNCH_OK("op_forward_kernel");
    op_stats_kernel<<<d->B, 32 * d->groups, 0, st>>>(w.part, sv.stats, nblk, d->groups, (double)(OP_F / d->groups) * d->N * d->k, d->eps);
    GCANET_LAUNCH_OK("op_stats_kernel");
    op_finish_kernel<<<dim3(ceil_div(d->N, 8), d->B), 256, 0, st>>>(a);
    GCANET_LAUNCH_OK("op_finish_kernel");
    return GCANET_OK;
}

extern "C" int gcanet_offset_pred_backward(const gcanet_offset_desc *d, const float *points, const float *feature, const float *inst,
                                           const int32_t *key_index, const float *conv_w, const float *gamma, const float *beta,
                                           const float *att_w1, const float *att_w2, const float *off_w, const float *off_b,
                                           const float *grad_out, const void *saved, float *grad_feature, float *grad_inst,
                                           float *grad_conv_w, float *grad_gamma, float *grad_beta, float *grad_att_w1,
                                           float *grad_att_w2, float *grad_off_w, float *grad_off_b, void *ws, size_t ws_bytes,
                                           gcanet_stream_t stream) {
    int rc = op_check(d);
    if (rc) return rc;
    GCANET_REQUIRE(points && feature && inst && key_index && conv_w && gamma && beta && att_w1 && att_w2 && off_w && off_b && grad_out &&
                   saved && grad_feature && grad_inst && grad_conv_w && grad_gamma && grad_beta && grad_att_w1 && grad_att_w2 &&
                   grad_off_w && grad_off_b, "offset_pred_backward: null pointer");
    const size_t need = op_plan_ws(d, nullptr, nullptr);
    if (ws == nullptr || ws_bytes < need || reinterpret_cast<uintptr_t>(ws) % kAlign) {
        set_error("offset_pred_backward: workspace too small or misaligned (%zu given, %zu needed)", ws_bytes, need);
        return GCANET_ERR_WORKSPACE;
    }
    OpSaved sv; OpWs w;
    op_plan_saved(d, const_cast<void *>(saved), &sv);
    op_plan_ws(d, ws, &w);
    cudaStream_t st = as_stream(stream);
    OpBwdArgs p{};
    p.f = op_args(d, points, feature, inst, key_index, conv_w, gamma, beta, att_w1, att_w2, off_w, off_b, sv, w.part, nullptr);
    p.gout = grad_out; p.du = w.du; p.rpart = w.rpart; p.owpart = w.owpart; p.coef = w.coef; p.dfeat = grad_feature; p.dinst = grad_inst;
    p.dT = w.dT; p.dkn = w.dkn; p.dw1 = w.dw; p.dw2 = w.dw + d->k * d->k; p.dwp = w.dwp;
    const size_t bs = (size_t)d->B * d->S;
    GCANET_CUDA_OK(cudaMemsetAsync(w.dT, 0, (bs * OP_F + bs * d->E + 2 * d->k * d->k + OP_F * 3) * sizeof(float), st));
    const int nblk = ceil_div(d->N, OP_PTS);
    op_bwd_reduce_kernel<<<dim3(nblk, d->B), OP_WARPS * 32, 0, st>>>(p);
    GCANET_LAUNCH_OK("op_bwd_reduce_kernel");
    op_bwd_coef_kernel<<<d->B, 256, 0, st>>>(p, grad_gamma, grad_beta, grad_off_w, grad_off_b, w.sbc, nblk);
    GCANET_LAUNCH_OK("op_bwd_coef_kernel");
    op_bwd_ow_kernel<<<ceil_div(6 * OP_F + 3, 32), dim3(32, 32), 0, st>>>(w.owpart, d->B * nblk, grad_off_w, grad_off_b);
    GCANET_LAUNCH_OK("op_bwd_ow_kernel");
    op_bwd_affine_kernel<<<1, OP_F, 0, st>>>(w.sbc, grad_gamma, grad_beta, d->B);
    GCANET_LAUNCH_OK("op_bwd_affine_kernel");
    const size_t smem = op_smem_floats(d->S, d->E, d->k) * sizeof(float);
    GCANET_REQUIRE(smem <= 227 * 1024, "offset_pred_backward: S=%d, E=%d need %zu bytes of shared memory", d->S, d->E, smem);
    GCANET_CUDA_OK(cudaFuncSetAttribute(op_bwd_main_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    op_bwd_main_kernel<<<dim3(ceil_div(d->N, OP_PTS_BWD), d->B), OP_WARPS * 32, smem, st>>>(p);
    GCANET_LAUNCH_OK("op_bwd_main_kernel");
    op_bwd_keys_kernel<<<dim3(d->S, d->B), 128, 0, st>>>(p);
    GCANET_LAUNCH_OK("op_bwd_keys_kernel");
    op_bwd_convw_kernel<<<dim3(OP_F / 8, d->B), OP_F + 3, 0, st>>>(p, w.cwpart);
    GCANET_LAUNCH_OK("op_bwd_convw_kernel");
    op_bwd_convw_sum_kernel<<<OP_F, OP_F + 3, 0, st>>>(p, w.cwpart, grad_conv_w);
    GCANET_LAUNCH_OK("op_bwd_convw_sum_kernel");
    GCANET_CUDA_OK(cudaMemcpyAsync(grad_att_w1, p.dw1, (size_t)d->k * d->k * sizeof(float), cudaMemcpyDeviceToDevice, st));
    GCANET_CUDA_OK(cudaMemcpyAsync(grad_att_w2, p.dw2, (size_t)d->k * d->k * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return GCANET_OK;
}
