// Library bookkeeping (status strings, per-thread error message, device check) and the
// channel-major <-> point-major transposes.
#include "common.cuh"

#include <string.h>
#include <atomic>

namespace gcanet {

static thread_local char g_err[512] = "";

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// Measurement probe (gcanet_knn_probe_arm / _read): per-thread pair of timing events around the kNN scan kernel(s).
struct ScanProbe {
    cudaEvent_t begin = nullptr, end = nullptr;
    bool armed = false, open = false, fired = false;
};
static thread_local ScanProbe g_probe;

void probe_scan_begin(cudaStream_t st) {
    if (!g_probe.armed || g_probe.open) return;
    if (cudaEventRecord(g_probe.begin, st) == cudaSuccess) g_probe.open = true;
}
void probe_scan_end(cudaStream_t st) {
    if (!g_probe.open) return;
    g_probe.fired = cudaEventRecord(g_probe.end, st) == cudaSuccess;
    g_probe.open = g_probe.armed = false;
}

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// x_cn [B][C][N] -> x_nc [B][N][ld]; columns C..ld-1 are zero-filled.  ADD: dst = transpose(src) + add, with
// add [B][N][ld] laid out like dst (the sum of a channel-major and a point-major incoming gradient in one pass).
template <bool ADD>
__global__ void cn_to_nc_kernel(const float *__restrict__ src, const float *__restrict__ add, float *__restrict__ dst, int C, int N, int ld) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const float *s = src + (size_t)b * C * N;
    float *d = dst + (size_t)b * N * ld;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int c = c0 + r, n = n0 + threadIdx.x;
        tile[r][threadIdx.x] = (c < C && n < N) ? s[(size_t)c * N + n] : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int n = n0 + r, c = c0 + threadIdx.x;
        if (n < N && c < ld) {
            float v = tile[threadIdx.x][r];
            if (ADD) v += __ldg(add + (size_t)b * N * ld + (size_t)n * ld + c);
            d[(size_t)n * ld + c] = v;
        }
    }
}

// Same, 64 x 64 tiles with 16-byte global accesses on both sides (N % 4 == 0, ld % 4 == 0, 16-byte aligned bases):
// four times the bytes per thread of the 32 x 32 kernel, which was latency-bound at ~55 % of the HBM rate.
template <bool ADD>
__global__ void __launch_bounds__(256) cn_to_nc_wide_kernel(const float *__restrict__ src, const float *__restrict__ add, float *__restrict__ dst, int C, int N, int ld) {
    __shared__ float tile[64][65];
    const int b = blockIdx.z;
    const int n0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
    const float *s = src + (size_t)b * C * N;
    float *d = dst + (size_t)b * N * ld;
    float4 a[4];
    if (ADD) {   // issued before the transposing loads so both streams are in flight together
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = threadIdx.x + 256 * i, n = e >> 4, c = (e & 15) * 4;
            a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + n < N && c0 + c < ld)
                a[i] = __ldg(reinterpret_cast<const float4 *>(add + (size_t)b * N * ld + (size_t)(n0 + n) * ld + c0 + c));
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int e = threadIdx.x + 256 * i, c = e >> 4, n = (e & 15) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c0 + c < C && n0 + n < N) v = __ldg(reinterpret_cast<const float4 *>(s + (size_t)(c0 + c) * N + n0 + n));
        tile[c][n] = v.x; tile[c][n + 1] = v.y; tile[c][n + 2] = v.z; tile[c][n + 3] = v.w;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int e = threadIdx.x + 256 * i, n = e >> 4, c = (e & 15) * 4;
        if (n0 + n < N && c0 + c < ld) {
            float4 v = make_float4(tile[c][n], tile[c + 1][n], tile[c + 2][n], tile[c + 3][n]);
            if (ADD) { v.x += a[i].x; v.y += a[i].y; v.z += a[i].z; v.w += a[i].w; }
            *reinterpret_cast<float4 *>(d + (size_t)(n0 + n) * ld + c0 + c) = v;
        }
    }
}

// x_nc [B][N][ld] -> x_cn [B][C][N]
__global__ void nc_to_cn_kernel(const float *__restrict__ src, float *__restrict__ dst, int C, int N, int ld) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const float *s = src + (size_t)b * N * ld;
    float *d = dst + (size_t)b * C * N;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int n = n0 + r, c = c0 + threadIdx.x;
        tile[r][threadIdx.x] = (n < N && c < C) ? s[(size_t)n * ld + c] : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int c = c0 + r, n = n0 + threadIdx.x;
        if (c < C && n < N) d[(size_t)c * N + n] = tile[threadIdx.x][r];
    }
}

int launch_cn_to_nc(const float *x_cn, float *x_nc, int B, int C, int N, int ld, cudaStream_t st, const float *add_nc) {
    const uintptr_t bases = reinterpret_cast<uintptr_t>(x_cn) | reinterpret_cast<uintptr_t>(x_nc) | reinterpret_cast<uintptr_t>(add_nc);
    if (ld >= 32 && N % 4 == 0 && ld % 4 == 0 && (bases & 15) == 0) {
        const dim3 grid(ceil_div(N, 64), ceil_div(ld, 64), B);
        if (add_nc) cn_to_nc_wide_kernel<true><<<grid, 256, 0, st>>>(x_cn, add_nc, x_nc, C, N, ld);
        else cn_to_nc_wide_kernel<false><<<grid, 256, 0, st>>>(x_cn, nullptr, x_nc, C, N, ld);
        GCANET_LAUNCH_OK("cn_to_nc_wide_kernel");
        return GCANET_OK;
    }
    dim3 grid(ceil_div(N, 32), ceil_div(ld, 32), B), block(32, 8);
    if (add_nc) cn_to_nc_kernel<true><<<grid, block, 0, st>>>(x_cn, add_nc, x_nc, C, N, ld);
    else cn_to_nc_kernel<false><<<grid, block, 0, st>>>(x_cn, nullptr, x_nc, C, N, ld);
    GCANET_LAUNCH_OK("cn_to_nc_kernel");
    return GCANET_OK;
}

int launch_nc_to_cn(const float *x_nc, float *x_cn, int B, int C, int N, int ld, cudaStream_t st) {
    dim3 grid(ceil_div(N, 32), ceil_div(C, 32), B), block(32, 8);
    nc_to_cn_kernel<<<grid, block, 0, st>>>(x_nc, x_cn, C, N, ld);
    GCANET_LAUNCH_OK("nc_to_cn_kernel");
    return GCANET_OK;
}

}  // namespace gcanet

using namespace gcanet;

extern "C" int gcanet_abi_version(void) { return GCANET_ABI_VERSION; }

extern "C" const char *gcanet_last_error(void) { return g_err; }

extern "C" unsigned long long gcanet_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int gcanet_knn_probe_arm(int on) {
    if (on && g_probe.begin == nullptr) {
        GCANET_CUDA_OK(cudaEventCreate(&g_probe.begin));
        GCANET_CUDA_OK(cudaEventCreate(&g_probe.end));
    }
    g_probe.armed = on != 0;
    g_probe.open = g_probe.fired = false;
    return GCANET_OK;
}

extern "C" int gcanet_knn_probe_read(float *scan_ms) {
    GCANET_REQUIRE(scan_ms != nullptr, "knn_probe_read: null pointer");
    GCANET_REQUIRE(g_probe.fired, "knn_probe_read: no tensor-core scan has run on this thread since the probe was armed");
    GCANET_CUDA_OK(cudaEventSynchronize(g_probe.end));
    GCANET_CUDA_OK(cudaEventElapsedTime(scan_ms, g_probe.begin, g_probe.end));
    return GCANET_OK;
}

extern "C" const char *gcanet_status_string(int status) {
    switch (status) {
        case GCANET_OK: return "ok";
        case GCANET_ERR_INVALID_ARGUMENT: return "invalid argument";
        case GCANET_ERR_WORKSPACE: return "workspace too small or misaligned";
        case GCANET_ERR_CUDA: return "CUDA error";
        case GCANET_ERR_UNSUPPORTED_DEVICE: return "unsupported device (needs sm_100)";
        default: return "unknown status";
    }
}

extern "C" int gcanet_check_device(void) {
    int dev = 0, major = 0;
    GCANET_CUDA_OK(cudaGetDevice(&dev));
    GCANET_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) {
        set_error("device %d has compute capability major %d; this library is built for sm_100a only", dev, major);
        return GCANET_ERR_UNSUPPORTED_DEVICE;
    }
    return GCANET_OK;
}

extern "C" int gcanet_cn_to_nc(const float *x_cn, float *x_nc, int B, int C, int N, int ld, gcanet_stream_t stream) {
    GCANET_REQUIRE(x_cn && x_nc, "cn_to_nc: null pointer");
    GCANET_REQUIRE(B >= 1 && C >= 1 && N >= 1 && ld >= C, "cn_to_nc: bad shape B=%d C=%d N=%d ld=%d", B, C, N, ld);
    GCANET_REQUIRE(B <= 65535, "cn_to_nc: B > 65535");
    return launch_cn_to_nc(x_cn, x_nc, B, C, N, ld, as_stream(stream));
}

extern "C" int gcanet_cn_to_nc_add(const float *x_cn, const float *add_nc, float *x_nc, int B, int C, int N, int ld,
                                   gcanet_stream_t stream) {
    GCANET_REQUIRE(x_cn && x_nc && add_nc, "cn_to_nc_add: null pointer");
    GCANET_REQUIRE(B >= 1 && C >= 1 && N >= 1 && ld >= C, "cn_to_nc_add: bad shape B=%d C=%d N=%d ld=%d", B, C, N, ld);
    GCANET_REQUIRE(B <= 65535, "cn_to_nc_add: B > 65535");
    return launch_cn_to_nc(x_cn, x_nc, B, C, N, ld, as_stream(stream), add_nc);
}

extern "C" int gcanet_nc_to_cn(const float *x_nc, float *x_cn, int B, int C, int N, int ld, gcanet_stream_t stream) {
    GCANET_REQUIRE(x_cn && x_nc, "nc_to_cn: null pointer");
    GCANET_REQUIRE(B >= 1 && C >= 1 && N >= 1 && ld >= C, "nc_to_cn: bad shape B=%d C=%d N=%d ld=%d", B, C, N, ld);
    GCANET_REQUIRE(B <= 65535, "nc_to_cn: B > 65535");
    return launch_nc_to_cn(x_nc, x_cn, B, C, N, ld, as_stream(stream));
}
