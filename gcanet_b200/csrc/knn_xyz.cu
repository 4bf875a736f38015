// Exact kNN for 3-D point clouds (xyz layer: L2 on C = 3, and the points x normals metric on
// C = 6) with spatial pruning.  Same result as the brute-force scan -- same fp32 distance
// arithmetic, ties resolved by (distance, original index) -- but only a fraction of the
// N x N pairs is evaluated:
//
//   1. per cloud: bounding box -> Morton code per point (up to 10 bits per axis) -> radix sort (CUB) of
//      (cloud, code) keys; coordinates, norms and the permutation are gathered into sorted
//      SoA arrays; every 32 consecutive sorted points form a tile with an AABB.
//   2. one CTA per query tile (8 warps x 4 queries).  A warp first scans its own tile and the
//      Morton-adjacent ones, which brings the per-query thresholds close to the final k-th
//      distance; then lanes test 32 tile AABBs at a time against the warp's query box and only
//      tiles whose lower bound is within the largest threshold are scanned (one reference per lane).
//      Candidate lists / thresholds work exactly as in knn_select.cu.
//
// Lower bound: for the L2 metric d >= dist^2(AABB_q, AABB_t); for the points x normals metric
// d = d_p * (1 + d_n) with 1 + d_n >= 3 - 2 max|n|^2 =: c (c = 1 for unit normals), so
// d >= c * d_p when c > 0; clouds with c <= 0 (normals far from unit length) take the brute-force
// path.  Bounds are shrunk by a relative 1e-5 and an absolute 1e-6 * extent^2 so fp32 rounding in the
// expansion-form distance can never prune a true neighbour.
#include "knn_lists.cuh"

#include <cub/device/device_radix_sort.cuh>

namespace gcanet {

constexpr int XW = 8;                 // warps per CTA
constexpr int XR = 4;                 // queries per warp
constexpr int XQ = XW * XR;           // = 32 = one tile of queries per CTA
constexpr int XT = 32;                // points per tile

// ---------------------------------------------------------------------------------
// prep
// ---------------------------------------------------------------------------------
// bbox[b] = (minx, miny, minz, maxx, maxy, maxz, max |n|^2, unused)
__global__ void xyz_bbox_kernel(const float *__restrict__ x, float *__restrict__ bbox, int C, int N) {
    __shared__ float red[7][32];
    const int b = blockIdx.x;
    const float *p = x + (size_t)b * C * N;
    float mn[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, mx[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    float nmax = 0.f;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v = p[(size_t)c * N + n];
            mn[c] = fminf(mn[c], v);
            mx[c] = fmaxf(mx[c], v);
        }
        if (C == 6) {
            float a0 = p[(size_t)3 * N + n], a1 = p[(size_t)4 * N + n], a2 = p[(size_t)5 * N + n];
            nmax = fmaxf(nmax, a0 * a0 + a1 * a1 + a2 * a2);
        }
    }
    float vals[7] = {mn[0], mn[1], mn[2], mx[0], mx[1], mx[2], nmax};
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        float v = vals[i];
        for (int o = 16; o; o >>= 1) {
            float w = __shfl_xor_sync(FULL, v, o);
            v = i < 3 ? fminf(v, w) : fmaxf(v, w);
        }
        if ((threadIdx.x & 31) == 0) red[i][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 7) {
        const int i = threadIdx.x;
        float v = red[i][0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) v = i < 3 ? fminf(v, red[i][w]) : fmaxf(v, red[i][w]);
        bbox[b * 8 + i] = v;
    }
}

// key = (cloud << 3*ab) | Hilbert-curve position (common.cuh) with `ab` bits per axis: fits a 32-bit radix key (four 8-bit passes at B = 16,
// ab = 9, instead of five over 64-bit keys); 512 cells per axis are far finer than the point spacing of a 10 k cloud,
// and the order only shapes the tiles -- exactness rests on their AABBs
__global__ void xyz_code_kernel(const float *__restrict__ x, const float *__restrict__ bbox,
                                unsigned *__restrict__ keys, int *__restrict__ vals, int C, int N, int ab) {
    const int b = blockIdx.y;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float *p = x + (size_t)b * C * N;
    const float *bb = bbox + b * 8;
    const float ext = fmaxf(fmaxf(bb[3] - bb[0], bb[4] - bb[1]), fmaxf(bb[5] - bb[2], 1e-30f));
    const float top = (float)((1 << ab) - 1);
    const float sc = top / ext;
    unsigned q[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float t = (p[(size_t)c * N + n] - bb[c]) * sc;
        q[c] = (unsigned)fminf(fmaxf(t, 0.f), top);
    }
    const unsigned code = hilbert3(q[0], q[1], q[2], ab);
    keys[(size_t)b * N + n] = ((unsigned)b << (3 * ab)) | code;
    vals[(size_t)b * N + n] = n;
}

// sorted SoA: sc[c][b][n] (c < C), sn[b][n] = |xyz|^2 in reference order, perm[b][n] = original index;
// aabb[b][t][6] for tiles of 32 sorted points
__global__ void xyz_gather_kernel(const float *__restrict__ x, const float *__restrict__ norm, const int *__restrict__ sorted_vals,
                                  float *__restrict__ sc, float *__restrict__ sn, int *__restrict__ perm,
                                  float *__restrict__ aabb, int B, int C, int N, int tiles) {
    const int b = blockIdx.y;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= tiles) return;
    const int n = warp * XT + lane;
    float v[3] = {0.f, 0.f, 0.f};
    float mn[3], mx[3];
    if (n < N) {
        const int o = sorted_vals[(size_t)b * N + n];
        perm[(size_t)b * N + n] = o;
        sn[(size_t)b * N + n] = norm[(size_t)b * N + o];
        for (int c = 0; c < C; ++c) {
            float t = x[((size_t)b * C + c) * N + o];
            sc[((size_t)c * B + b) * N + n] = t;
            if (c < 3) v[c] = t;
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        mn[c] = n < N ? v[c] : CUDART_INF_F;
        mx[c] = n < N ? v[c] : -CUDART_INF_F;
        for (int o = 16; o; o >>= 1) {
            mn[c] = fminf(mn[c], __shfl_xor_sync(FULL, mn[c], o));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(FULL, mx[c], o));
        }
    }
    if (lane == 0) {
        float *a = aabb + ((size_t)b * tiles + warp) * 6;
        a[0] = mn[0]; a[1] = mn[1]; a[2] = mn[2]; a[3] = mx[0]; a[4] = mx[1]; a[5] = mx[2];
    }
}

// ---------------------------------------------------------------------------------
// pruned scan
// ---------------------------------------------------------------------------------
struct XyzArgs {
    const float *sc;      // [C][B][N] sorted coordinates
    const float *sn;      // [B][N]    sorted xyz norms
    const int *perm;      // [B][N]
    const float *aabb;    // [B][tiles][6]
    const float *bbox;    // [B][8]
    int B, N, tiles, k, step, kout;
    int64_t *idx64;
    int32_t *idx32;
    int *fallback;        // [B] set to 1 when the cloud must take the brute-force path (PN, c <= 0)
    int unordered;        // 1: only the neighbour set is needed, skip the final ordering
    int aabb_smem;        // 1: the cloud's tile boxes are copied to shared memory; 0: read through L1 (large clouds, where
                          //    the table would cost the CTA its co-residents)
};

template <int CDIM, bool PN, int SL>
__global__ void __launch_bounds__(XW * 32, 3) knn_xyz_kernel(XyzArgs a) {   // 80 registers: three CTAs (24 warps) per SM hide the list latency better than two
    extern __shared__ __align__(16) float smem[];
    constexpr int CAP = 32 * SL;
    float *s_aabb = smem;                                     // [tiles][6] (aabb_smem only)
    float *lds = s_aabb + (a.aabb_smem ? (size_t)a.tiles * 6 : 0);   // [XQ][CAP]
    int *lis = reinterpret_cast<int *>(lds + XQ * CAP);       // [XQ][CAP]

    const int b = blockIdx.y;
    const int tile0 = blockIdx.x;                             // this CTA's query tile
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int N = a.N, tiles = a.tiles;
    const float *bb = a.bbox + b * 8;

    float cmin = 1.f;                                          // d >= cmin * d_p
    if (PN) {
        cmin = 3.f - 2.f * bb[6];
        if (cmin <= 0.05f) {                                   // normals far from unit: no usable bound
            if (threadIdx.x == 0 && blockIdx.x == 0) a.fallback[b] = 1;
            return;
        }
        cmin = fminf(cmin, 1.f);
    }
    const float ext = fmaxf(fmaxf(bb[3] - bb[0], bb[4] - bb[1]), bb[5] - bb[2]);
    const float abs_slack = 1e-6f * ext * ext + 2e-5f * fmaxf(fmaxf(bb[3] * bb[3], bb[0] * bb[0]),
                                                              fmaxf(fmaxf(bb[4] * bb[4], bb[1] * bb[1]), fmaxf(bb[5] * bb[5], bb[2] * bb[2])));

    const float *boxes = a.aabb + (size_t)b * tiles * 6;
    if (a.aabb_smem) {
        for (int e = threadIdx.x; e < tiles * 6; e += XW * 32) s_aabb[e] = boxes[e];
        __syncthreads();
        boxes = s_aabb;
    }

    // my queries: sorted positions q0 .. q0+3
    const int q0 = tile0 * XT + warp * XR;
    float qc[XR][CDIM], qn[XR];
    float qmin[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, qmax[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
#pragma unroll
    for (int r = 0; r < XR; ++r) {
        const int q = min(q0 + r, N - 1);
#pragma unroll
        for (int c = 0; c < CDIM; ++c) qc[r][c] = a.sc[((size_t)c * a.B + b) * N + q];
        qn[r] = a.sn[(size_t)b * N + q];
#pragma unroll
        for (int c = 0; c < 3; ++c) { qmin[c] = fminf(qmin[c], qc[r][c]); qmax[c] = fmaxf(qmax[c], qc[r][c]); }
    }

    float thr[XR];
    int cnt[XR];
#pragma unroll
    for (int r = 0; r < XR; ++r) { thr[r] = CUDART_INF_F; cnt[r] = 0; }
    float *my_ld = lds + (size_t)(warp * XR) * CAP;
    int *my_li = lis + (size_t)(warp * XR) * CAP;

    auto scan_tile = [&](int t) {
        const int j = t * XT + lane;                           // sorted position of my reference
        const bool valid = j < N;
        const int jj = valid ? j : N - 1;
        float rv[CDIM];
#pragma unroll
        for (int c = 0; c < CDIM; ++c) rv[c] = __ldg(a.sc + ((size_t)c * a.B + b) * N + jj);
        const float rnorm = __ldg(a.sn + (size_t)b * N + jj);
        const int orig = __ldg(a.perm + (size_t)b * N + jj);
        static_for<0, XR>([&](auto rc) {
            constexpr int r = decltype(rc)::value;
            float tp = __fmul_rn(qc[r][0], rv[0]);
            tp = fmaf(qc[r][1], rv[1], tp);
            tp = fmaf(qc[r][2], rv[2], tp);
            float d = __fadd_rn(fmaf(-2.f, tp, rnorm), qn[r]);
            if constexpr (PN) {
                float tn = __fmul_rn(qc[r][3 % CDIM], rv[3 % CDIM]);
                tn = fmaf(qc[r][4 % CDIM], rv[4 % CDIM], tn);
                tn = fmaf(qc[r][5 % CDIM], rv[5 % CDIM], tn);
                d = __fmul_rn(d, __fadd_rn(1.f, fmaf(-2.f, tn, 2.f)));
            }
            // visiting order is not index order: accept ties with the threshold, (distance, index) ranking decides
            const bool pass = valid && d <= thr[r];
            const unsigned m = __ballot_sync(FULL, pass);
            if (m) {
                if (pass) {
                    const int p = cnt[r] + __popc(m & ((1u << lane) - 1));
                    my_ld[r * CAP + p] = d;
                    my_li[r * CAP + p] = orig;
                }
                cnt[r] += __popc(m);
                if (cnt[r] > CAP - 32) {
                    __syncwarp();
                    float t = thr[r];          // by-reference argument of a non-inlined call: keep thr[] itself in registers
                    cnt[r] = shrink_list<SL>(my_ld + r * CAP, my_li + r * CAP, cnt[r], a.k, CAP - 32, lane, t);
                    thr[r] = t;
                }
            }
        });
    };

    // Two passes over the tile list with ONE copy of the scan code:
    //   pass 0: own tile and its Morton neighbours, unconditionally -- brings the thresholds close to
    //           the final k-th distances before any pruning decision (the window is wide enough for
    //           every query to see at least k + 32 references);
    //   pass 1: every other tile, only if its AABB lower bound is within the largest threshold.
    int lo_t, hi_t;
    {
        int half = (a.k + XT - 1) / XT;
        lo_t = max(0, tile0 - half);
        hi_t = min(tiles - 1, tile0 + half);
        while ((hi_t - lo_t + 1) * XT < a.k + 2 * XT && (lo_t > 0 || hi_t < tiles - 1)) {
            if (lo_t > 0) --lo_t;
            if (hi_t < tiles - 1) ++hi_t;
        }
    }
    for (int pass = 0; pass < 2; ++pass) {
        const int b0 = pass == 0 ? (lo_t / 32) * 32 : 0;
        const int b1 = pass == 0 ? hi_t + 1 : tiles;
        for (int base = b0; base < b1; base += 32) {
            const int t = base + lane;
            const bool in_window = t >= lo_t && t <= hi_t;
            const bool mine = t < tiles && (pass == 0 ? in_window : !in_window);
            float lb = -CUDART_INF_F;                                  // pass 0: always scanned
            if (mine && pass == 1) {
                const float *bx = boxes + t * 6;
                float acc = 0.f;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float g = fmaxf(fmaxf(bx[c] - qmax[c], qmin[c] - bx[3 + c]), 0.f);
                    acc = fmaf(g, g, acc);
                }
                lb = cmin * acc * (1.f - 1e-5f) - abs_slack;
            }
            float tmax = fmaxf(fmaxf(thr[0], thr[1]), fmaxf(thr[2], thr[3]));
            unsigned todo = __ballot_sync(FULL, mine && lb <= tmax);
            while (todo) {
                const int l = __ffs(todo) - 1;
                todo &= todo - 1;
                const float lbt = __shfl_sync(FULL, lb, l);
                tmax = fmaxf(fmaxf(thr[0], thr[1]), fmaxf(thr[2], thr[3]));
                if (lbt <= tmax) scan_tile(base + l);
            }
        }
    }

    // rank the survivors and write the k best in order, at the ORIGINAL query positions
    __syncwarp();
    static_for<0, XR>([&](auto rc) {
        constexpr int r = decltype(rc)::value;
        if (q0 + r >= N) return;
        float t = thr[r];
        int n = cnt[r];
        float *ld = my_ld + r * CAP;
        int *li = my_li + r * CAP;
        if (a.unordered) {
            if (n > a.k) select_k_unordered<SL>(ld, li, n, a.k, lane);                    // k best, any order
        } else {
            if (n > a.k + kSlack) n = shrink_list<SL>(ld, li, n, a.k, CAP, lane, t);
            rank_cut<SL>(ld, li, n, a.k, lane);                                           // k best, in order
        }
    });
    __syncwarp();
#pragma unroll
    for (int r = 0; r < XR; ++r) {
        if (q0 + r >= N) continue;
        const int qo = a.perm[(size_t)b * N + q0 + r];
        for (int p = lane; p < a.k; p += 32) {
            if (p % a.step) continue;
            const size_t o = ((size_t)b * N + qo) * a.kout + p / a.step;
            const int id = my_li[r * CAP + p];
            if (a.idx64) a.idx64[o] = id;
            if (a.idx32) a.idx32[o] = id;
        }
    }
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
static size_t cub_temp_bytes(size_t n, int end_bit) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned *)nullptr, (unsigned *)nullptr,
                                    (const int *)nullptr, (int *)nullptr, (int)n, 0, end_bit);
    return bytes;
}

static int cloud_bits(int B) {
    int bits = 0;
    while ((1 << bits) < B) ++bits;
    return bits;
}
static int axis_bits(int B) {                 // B <= 65535 (checked by the caller): at least 5 bits per axis
    const int ab = (32 - cloud_bits(B)) / 3;
    return ab > 10 ? 10 : ab;
}
static int key_bits(int B) { return cloud_bits(B) + 3 * axis_bits(B); }

bool knn_xyz_supported(int C, int N, int k2, int metric) {
    if (!((metric == GCANET_METRIC_L2 && C == 3) || (metric == GCANET_METRIC_POINTS_NORMALS && C == 6))) return false;
    return N >= 256 && k2 <= 168 && (size_t)((N + XT - 1) / XT) * 24 <= 96 * 1024;
}

size_t knn_xyz_workspace_bytes(int B, int C, int N) {
    const size_t bn = (size_t)B * N;
    const int tiles = ceil_div(N, XT);
    size_t t = 0;
    t += align_up(bn * sizeof(float));                     // norm (reference order, original positions)
    t += align_up((size_t)B * 8 * sizeof(float));          // bbox
    t += 2 * align_up(bn * sizeof(unsigned));              // keys in/out
    t += 2 * align_up(bn * sizeof(int));                   // vals in/out
    t += align_up(cub_temp_bytes(bn, key_bits(B)));        // cub temp
    t += align_up(bn * C * sizeof(float));                 // sorted coords
    t += align_up(bn * sizeof(float));                     // sorted norms
    t += align_up(bn * sizeof(int));                       // perm
    t += align_up((size_t)B * tiles * 6 * sizeof(float));  // aabb
    t += align_up((size_t)B * sizeof(int));                // fallback flags
    return t;
}

// declared in knn_select.cu
int launch_sqnorm_public(const float *x, float *out, int B, int C, int Cuse, int N, cudaStream_t st);

template <int CDIM, bool PN>
static int launch_xyz(XyzArgs a, cudaStream_t st) {
    dim3 grid(a.tiles, a.B);
    auto go = [&](auto slc) -> int {
        constexpr int SL = decltype(slc)::value;
        size_t smem = ((a.aabb_smem ? (size_t)a.tiles * 6 : 0) + 2 * (size_t)XQ * 32 * SL) * sizeof(float);
        auto kern = knn_xyz_kernel<CDIM, PN, SL>;
        if (smem > 48 * 1024) GCANET_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, XW * 32, smem, st>>>(a);
        GCANET_LAUNCH_OK("knn_xyz_kernel");
        return GCANET_OK;
    };
    // list capacity 32 * SL per query: a shrink must leave room for a tile (k + 8 <= 32 (SL - 1)); the roomier choice
    // (fewer shrinks) measures best
    const int sl = a.k <= 40 ? 4 : (a.k <= 72 ? 5 : (a.k <= 104 ? 6 : 8));
    if (sl == 4) return go(std::integral_constant<int, 4>{});
    if (sl == 5) return go(std::integral_constant<int, 5>{});
    if (sl == 6) return go(std::integral_constant<int, 6>{});
    return go(std::integral_constant<int, 8>{});
}

// the sort keys alone, for the tensor-core scan of xyz clouds (knn_tc.cu)
int launch_xyz_sort_keys(const float *x, float *bbox, unsigned *keys, int *vals, int B, int C, int N, int ab, cudaStream_t st) {
    xyz_bbox_kernel<<<B, 1024, 0, st>>>(x, bbox, C, N);
    GCANET_LAUNCH_OK("xyz_bbox_kernel");
    xyz_code_kernel<<<dim3(ceil_div(N, 256), B), 256, 0, st>>>(x, bbox, keys, vals, C, N, ab);
    GCANET_LAUNCH_OK("xyz_code_kernel");
    return GCANET_OK;
}

// Returns GCANET_OK; *fallback_flags (device, [B]) tells the caller which clouds need the brute-force scan.
int knn_graph_xyz(const float *x, int B, int C, int N, int k1, int k2, int metric, int64_t *idx64, int32_t *idx32,
                  void *ws, float **norm_out, int **fallback_out, int unordered, cudaStream_t st) {
    const size_t bn = (size_t)B * N;
    const int tiles = ceil_div(N, XT);
    const int end_bit = key_bits(B);
    Carver cv(ws);
    float *norm = cv.take<float>(bn);
    float *bbox = cv.take<float>((size_t)B * 8);
    unsigned *keys_in = cv.take<unsigned>(bn);
    unsigned *keys_out = cv.take<unsigned>(bn);
    int *vals_in = cv.take<int>(bn);
    int *vals_out = cv.take<int>(bn);
    size_t temp_bytes = cub_temp_bytes(bn, end_bit);
    void *temp = cv.take<char>(temp_bytes);
    float *sc = cv.take<float>(bn * C);
    float *sn = cv.take<float>(bn);
    int *perm = cv.take<int>(bn);
    float *aabb = cv.take<float>((size_t)B * tiles * 6);
    int *fallback = cv.take<int>(B);

    int rc = launch_sqnorm_public(x, norm, B, C, 3, N, st);
    if (rc) return rc;
    GCANET_CUDA_OK(cudaMemsetAsync(fallback, 0, B * sizeof(int), st));
    xyz_bbox_kernel<<<B, 1024, 0, st>>>(x, bbox, C, N);
    GCANET_LAUNCH_OK("xyz_bbox_kernel");
    xyz_code_kernel<<<dim3(ceil_div(N, 256), B), 256, 0, st>>>(x, bbox, keys_in, vals_in, C, N, axis_bits(B));
    GCANET_LAUNCH_OK("xyz_code_kernel");
    GCANET_CUDA_OK(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, vals_in, vals_out, (int)bn, 0, end_bit, st));
    count_launch();
    xyz_gather_kernel<<<dim3(ceil_div(tiles, 8), B), 256, 0, st>>>(x, norm, vals_out, sc, sn, perm, aabb, B, C, N, tiles);
    GCANET_LAUNCH_OK("xyz_gather_kernel");

    XyzArgs a{sc, sn, perm, aabb, bbox, B, N, tiles, k2, k2 / k1, gcanet_knn_graph_columns(k1, k2), idx64, idx32, fallback,
              (unordered && k1 == k2) ? 1 : 0, tiles * 24 <= 16 * 1024 ? 1 : 0};
    rc = metric == GCANET_METRIC_L2 ? launch_xyz<3, false>(a, st) : launch_xyz<6, true>(a, st);
    if (rc) return rc;
    *norm_out = norm;
    *fallback_out = fallback;
    return GCANET_OK;
}

}  // namespace gcanet
