// Micro-benchmark of the two memory patterns the EdgeConv kernels are bound by:
//   (1) random-row GATHER out of an L2-resident buffer: one warp per destination point reads k rows of
//       row_bytes (128 / 256 / 512 B) at random positions of its cloud's [N][stride] matrix, eight rows in
//       flight per warp, one vector load per lane -- the access pattern of edge_gather_reduce_kernel;
//   (2) random-row vector REDUCTION: one warp per source point adds a row of row_bytes to k random rows of
//       its cloud with red.global.add.v2/v4.f32 -- the pattern of edge_bwd_scatter_kernel.
// Same geometry as the benchmark step: B = 16 clouds x N = 10 000 points x k = 50 edges per point, indices
// uniform inside the cloud (no L1 reuse: a conservative peak).  Prints one JSON object per configuration;
// "gbps" = edges * row_bytes / time.  These are the denominators profiles/ quotes the gather and the scatter
// against.   Build + run:  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/l2_peaks tools/l2_peaks.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int kWarps = 8, kPtsPerWarp = 4;

template <int VEC>
struct V;
template <> struct V<1> { using T = float;  };
template <> struct V<2> { using T = float2; };
template <> struct V<4> { using T = float4; };

__device__ __forceinline__ float sum(float v) { return v; }
__device__ __forceinline__ float sum(float2 v) { return v.x + v.y; }
__device__ __forceinline__ float sum(float4 v) { return v.x + v.y + v.z + v.w; }

template <int VEC>
__global__ void __launch_bounds__(kWarps * 32) gather_kernel(const float *__restrict__ buf, const int *__restrict__ idx,
                                                             float *__restrict__ out, int N, int k, int stride) {
    using T = typename V<VEC>::T;
    const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float *base = buf + (size_t)b * N * stride + lane * VEC;
    for (int pi = 0; pi < kPtsPerWarp; ++pi) {
        const int i = (blockIdx.x * kWarps + warp) * kPtsPerWarp + pi;
        if (i >= N) break;
        const int *ip = idx + ((size_t)b * N + i) * k;
        float acc = 0.f;
        for (int t0 = 0; t0 < k; t0 += 32) {
            const int cnt = min(32, k - t0);
            const int myj = lane < cnt ? ip[t0 + lane] : 0;
            int t = 0;
            for (; t + 8 <= cnt; t += 8) {
                T p[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int j = __shfl_sync(0xffffffffu, myj, t + u);
                    p[u] = __ldg(reinterpret_cast<const T *>(base + (unsigned)j * (unsigned)stride));
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) acc += sum(p[u]);
            }
            for (; t < cnt; ++t) {
                const int j = __shfl_sync(0xffffffffu, myj, t);
                acc += sum(__ldg(reinterpret_cast<const T *>(base + (unsigned)j * (unsigned)stride)));
            }
        }
        out[((size_t)b * N + i) * 32 + lane] = acc;
    }
}

__device__ __forceinline__ void red(float *p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void red(float *p, float2 v) { atomicAdd(reinterpret_cast<float2 *>(p), v); }
__device__ __forceinline__ void red(float *p, float4 v) { atomicAdd(reinterpret_cast<float4 *>(p), v); }

template <int VEC>
__global__ void __launch_bounds__(kWarps * 32) scatter_kernel(float *__restrict__ buf, const int *__restrict__ idx, int N, int k,
                                                              int stride) {
    using T = typename V<VEC>::T;
    const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *base = buf + (size_t)b * N * stride + lane * VEC;
    T val;
    float *vp = reinterpret_cast<float *>(&val);
    for (int v = 0; v < VEC; ++v) vp[v] = 1e-3f * (lane + v);
    for (int pi = 0; pi < kPtsPerWarp; ++pi) {
        const int i = (blockIdx.x * kWarps + warp) * kPtsPerWarp + pi;
        if (i >= N) break;
        const int *ip = idx + ((size_t)b * N + i) * k;
        for (int t0 = 0; t0 < k; t0 += 32) {
            const int cnt = min(32, k - t0);
            const int myj = lane < cnt ? ip[t0 + lane] : 0;
            for (int t = 0; t < cnt; ++t) {
                const int j = __shfl_sync(0xffffffffu, myj, t);
                red(base + (unsigned)j * (unsigned)stride, val);
            }
        }
    }
}

template <typename F>
static float time_ms(F launch, int reps) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) launch();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms / reps;
}

int main() {
    const int B = 16, N = 10000, k = 50, reps = 20;
    const size_t edges = (size_t)B * N * k;
    std::vector<int> h(edges);
    unsigned s = 12345u;
    for (size_t e = 0; e < edges; ++e) { s = s * 1664525u + 1013904223u; h[e] = (int)((s >> 8) % N); }
    int *idx;
    CK(cudaMalloc(&idx, edges * sizeof(int)));
    CK(cudaMemcpy(idx, h.data(), edges * sizeof(int), cudaMemcpyHostToDevice));
    float *buf, *out;
    const size_t buf_floats = (size_t)B * N * 256;                // up to stride 256 floats (1 KB) per row
    CK(cudaMalloc(&buf, buf_floats * sizeof(float)));
    CK(cudaMemset(buf, 0, buf_floats * sizeof(float)));
    CK(cudaMalloc(&out, (size_t)B * N * 32 * sizeof(float)));
    const dim3 grid((N + kWarps * kPtsPerWarp - 1) / (kWarps * kPtsPerWarp), B), block(kWarps * 32);
    struct Cfg { int vec, stride; };
    const Cfg cfgs[] = {{1, 32}, {1, 64}, {2, 64}, {2, 128}, {4, 128}, {4, 256}};
    for (const Cfg &c : cfgs) {
        const int row_bytes = c.vec * 32 * 4;
        auto g = [&]() {
            if (c.vec == 1) gather_kernel<1><<<grid, block>>>(buf, idx, out, N, k, c.stride);
            else if (c.vec == 2) gather_kernel<2><<<grid, block>>>(buf, idx, out, N, k, c.stride);
            else gather_kernel<4><<<grid, block>>>(buf, idx, out, N, k, c.stride);
        };
        auto r = [&]() {
            if (c.vec == 1) scatter_kernel<1><<<grid, block>>>(buf, idx, N, k, c.stride);
            else if (c.vec == 2) scatter_kernel<2><<<grid, block>>>(buf, idx, N, k, c.stride);
            else scatter_kernel<4><<<grid, block>>>(buf, idx, N, k, c.stride);
        };
        const float mg = time_ms(g, reps), mr = time_ms(r, reps);
        CK(cudaGetLastError());
        const double bytes = (double)edges * row_bytes;
        printf("{\"pattern\": \"gather\", \"row_bytes\": %d, \"row_stride_bytes\": %d, \"buffer_mb\": %.1f, \"edges\": %zu, \"ms\": %.4f, \"gbps\": %.1f}\n",
               row_bytes, c.stride * 4, (double)B * N * c.stride * 4 / 1e6, edges, mg, bytes / mg / 1e6);
        printf("{\"pattern\": \"red.add.v%d.f32\", \"row_bytes\": %d, \"row_stride_bytes\": %d, \"buffer_mb\": %.1f, \"edges\": %zu, \"ms\": %.4f, \"gbps\": %.1f}\n",
               c.vec, row_bytes, c.stride * 4, (double)B * N * c.stride * 4 / 1e6, edges, mr, bytes / mr / 1e6);
    }
    return 0;
}
