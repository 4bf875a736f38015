// Warp-maintained candidate lists in shared memory, shared by the CUDA-core kNN kernels
// (knn_select.cu: brute-force scan; knn_xyz.cu: spatially pruned scan).
#pragma once

#include "common.cuh"

#include <math_constants.h>
#include <type_traits>

namespace gcanet {

constexpr unsigned FULL = 0xffffffffu;

// ---------------------------------------------------------------------------------
// per-query candidate list in shared memory, maintained by one warp
// ---------------------------------------------------------------------------------
constexpr int kSlack = 8;     // the bisection stops once the bound keeps <= k + kSlack entries

// compile-time loop: keeps per-query state indexed by constants so it stays in registers
template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F &&f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

// Ranks the n entries of a list by (distance, index), keeps the min(n, k) smallest IN ORDER
// (entry of rank r moves to slot r) and returns the largest kept distance.
// (not inlined: called from rare paths of kernels that are otherwise unrolled over several queries;
// inlining every copy blows the instruction cache)
template <int SL>
__device__ __noinline__ float rank_cut(float *ld, int *li, int n, int k, int lane) {
    float dv[SL];
    int di[SL], rank[SL];
#pragma unroll
    for (int s = 0; s < SL; ++s) {
        const int e = s * 32 + lane;
        dv[s] = e < n ? ld[e] : CUDART_INF_F;
        di[s] = e < n ? li[e] : 0x7fffffff;
        rank[s] = 0;
    }
    for (int e = 0; e < n; ++e) {
        const float od = ld[e];
        const int oi = li[e];
#pragma unroll
        for (int s = 0; s < SL; ++s) {
            if (s * 32 >= n) break;
            rank[s] += (od < dv[s] || (od == dv[s] && oi < di[s])) ? 1 : 0;
        }
    }
    __syncwarp();
    float kth = -CUDART_INF_F;
#pragma unroll
    for (int s = 0; s < SL; ++s) {
        if (s * 32 + lane < n && rank[s] < k) { ld[rank[s]] = dv[s]; li[rank[s]] = di[s]; kth = fmaxf(kth, dv[s]); }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) kth = fmaxf(kth, __shfl_xor_sync(FULL, kth, o));
    __syncwarp();
    return kth;
}

// Set-only finish: removes the n - k largest entries by (distance, index) -- n - k <= kSlack after a
// shrink -- leaving the k smallest in slots [0, k) in arbitrary order.
template <int SL>
__device__ __noinline__ void drop_largest(float *ld, int *li, int n, int k, int lane) {
    while (n > k) {
        // lexicographic max of (d, idx) over the list
        float bd = -CUDART_INF_F;
        int bi = -1, bp = -1;
        for (int e = lane; e < n; e += 32) {
            const float d = ld[e];
            const int i = li[e];
            if (d > bd || (d == bd && i > bi)) { bd = d; bi = i; bp = e; }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const float od = __shfl_xor_sync(FULL, bd, o);
            const int oi = __shfl_xor_sync(FULL, bi, o);
            const int op = __shfl_xor_sync(FULL, bp, o);
            if (od > bd || (od == bd && oi > bi)) { bd = od; bi = oi; bp = op; }
        }
        // move the last entry into the hole
        if (lane == 0 && bp != n - 1) { ld[bp] = ld[n - 1]; li[bp] = li[n - 1]; }
        --n;
        __syncwarp();
    }
}

// Set-only finish for n > k entries: bisection on the distance until exactly k entries lie at or below the bound
// (the usual case: one pass of compaction and no ranking at all); if ties at the k-th distance make that impossible
// the tie group is trimmed by (distance, index) with drop_largest.  Leaves the k smallest in slots [0, k).
template <int SL>
__device__ __noinline__ void select_k_unordered(float *ld, int *li, int n, int k, int lane) {
    float dv[SL];
    int di[SL];
    float mn = CUDART_INF_F, mx = -CUDART_INF_F;
#pragma unroll
    for (int s = 0; s < SL; ++s) {
        const int e = s * 32 + lane;
        dv[s] = CUDART_INF_F;
        di[s] = 0x7fffffff;
        if (e < n) { dv[s] = ld[e]; di[s] = li[e]; mn = fminf(mn, dv[s]); mx = fmaxf(mx, dv[s]); }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(FULL, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
    }
    float lo = mn, hi = mx;
    int c_hi = n;
    for (int it = 0; it < 48 && c_hi > k; ++it) {
        const float mid = 0.5f * lo + 0.5f * hi;
        if (!(mid > lo && mid < hi)) break;            // interval exhausted: ties at hi
        int c = 0;
#pragma unroll
        for (int s = 0; s < SL; ++s) c += (dv[s] <= mid) ? 1 : 0;
        c = __reduce_add_sync(FULL, c);
        if (c >= k) { hi = mid; c_hi = c; } else { lo = mid; }
    }
    __syncwarp();
    int base = 0;
#pragma unroll
    for (int s = 0; s < SL; ++s) {
        const bool keep = dv[s] <= hi && s * 32 + lane < n;
        const unsigned m = __ballot_sync(FULL, keep);
        if (keep) { const int p = base + __popc(m & ((1u << lane) - 1)); ld[p] = dv[s]; li[p] = di[s]; }
        base += __popc(m);
    }
    __syncwarp();
    if (base > k) drop_largest<SL>(ld, li, base, k, lane);
}

// Shrinks a list of n > k entries to those with d <= bound, where count(d <= bound) >= k and,
// ties permitting, <= k + kSlack (bisection on the values).  If ties would keep more than
// `limit` entries the list is cut to exactly the k smallest by (distance, index) instead.
// Returns the new count; `thr` becomes the bound: later candidates must be strictly below it
// (they have larger indices, so an equal distance loses the tie anyway).
template <int SL>
__device__ __noinline__ int shrink_list(float *ld, int *li, int n, int k, int limit, int lane, float &thr) {
    float dv[SL];
    int di[SL];
    float mn = CUDART_INF_F, mx = -CUDART_INF_F;
#pragma unroll
    for (int s = 0; s < SL; ++s) {
        const int e = s * 32 + lane;
        dv[s] = CUDART_INF_F;
        di[s] = 0x7fffffff;
        if (e < n) { dv[s] = ld[e]; di[s] = li[e]; mn = fminf(mn, dv[s]); mx = fmaxf(mx, dv[s]); }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(FULL, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
    }
    float lo = mn, hi = mx;
    int c_hi = n;
    for (int it = 0; it < 32 && c_hi > k + kSlack; ++it) {
        const float mid = 0.5f * lo + 0.5f * hi;
        if (!(mid > lo && mid < hi)) break;            // interval exhausted: ties at hi
        int c = 0;
#pragma unroll
        for (int s = 0; s < SL; ++s) c += (dv[s] <= mid) ? 1 : 0;
        c = __reduce_add_sync(FULL, c);
        if (c >= k) { hi = mid; c_hi = c; } else { lo = mid; }
    }
    __syncwarp();
    if (c_hi > limit) {
        thr = rank_cut<SL>(ld, li, n, k, lane);
        return k;
    }
    int base = 0;
#pragma unroll
    for (int s = 0; s < SL; ++s) {
        const bool keep = dv[s] <= hi && s * 32 + lane < n;
        const unsigned m = __ballot_sync(FULL, keep);
        if (keep) { const int p = base + __popc(m & ((1u << lane) - 1)); ld[p] = dv[s]; li[p] = di[s]; }
        base += __popc(m);
    }
    __syncwarp();
    thr = hi;
    return base;
}

}  // namespace gcanet
