// Library bookkeeping (status strings, per-thread error message, device check) and the
// channel-major <-> point-major transposes.
#include "common.cuh"

#include <string.h>
#include <atomic>

namespace gcanet {

static thread_local char g_err[512] = "";

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// x_cn [B][C][N] -> x_nc [B][N][ld]; columns C..ld-1 are zero-filled.
__global__ void cn_to_nc_kernel(const float *__restrict__ src, float *__restrict__ dst, int C, int N, int ld) {
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
        if (n < N && c < ld) d[(size_t)n * ld + c] = tile[threadIdx.x][r];
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

int launch_cn_to_nc(const float *x_cn, float *x_nc, int B, int C, int N, int ld, cudaStream_t st) {
    dim3 grid(ceil_div(N, 32), ceil_div(ld, 32), B), block(32, 8);
    cn_to_nc_kernel<<<grid, block, 0, st>>>(x_cn, x_nc, C, N, ld);
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

extern "C" int gcanet_nc_to_cn(const float *x_nc, float *x_cn, int B, int C, int N, int ld, gcanet_stream_t stream) {
    GCANET_REQUIRE(x_cn && x_nc, "nc_to_cn: null pointer");
    GCANET_REQUIRE(B >= 1 && C >= 1 && N >= 1 && ld >= C, "nc_to_cn: bad shape B=%d C=%d N=%d ld=%d", B, C, N, ld);
    GCANET_REQUIRE(B <= 65535, "nc_to_cn: B > 65535");
    return launch_nc_to_cn(x_nc, x_cn, B, C, N, ld, as_stream(stream));
}
