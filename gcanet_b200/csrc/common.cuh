// Shared helpers for the gcanet_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/gcanet_b200.h"

namespace gcanet {

void set_error(const char *fmt, ...);
void count_launch();   // statistics only: bumps the counter gcanet_launch_count() reports
// measurement probe (gcanet_knn_probe_arm): no-ops unless the calling thread armed it; the first begin / end pair after
// arming records the thread's two timing events on `st`
void probe_scan_begin(cudaStream_t st);
void probe_scan_end(cudaStream_t st);

#define GCANET_REQUIRE(cond, ...)                         \
    do {                                                  \
        if (!(cond)) {                                    \
            ::gcanet::set_error(__VA_ARGS__);             \
            return GCANET_ERR_INVALID_ARGUMENT;           \
        }                                                 \
    } while (0)

#define GCANET_CUDA_OK(expr)                                                            \
    do {                                                                                \
        cudaError_t err__ = (expr);                                                     \
        if (err__ != cudaSuccess) {                                                     \
            ::gcanet::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), \
                                __FILE__, __LINE__);                                    \
            return GCANET_ERR_CUDA;                                                     \
        }                                                                               \
    } while (0)

#define GCANET_LAUNCH_OK(name)                                                          \
    do {                                                                                \
        cudaError_t err__ = cudaGetLastError();                                         \
        ::gcanet::count_launch();                                                       \
        if (err__ != cudaSuccess) {                                                     \
            ::gcanet::set_error("launch of %s failed: %s", name, cudaGetErrorString(err__)); \
            return GCANET_ERR_CUDA;                                                     \
        }                                                                               \
    } while (0)

// Measurement knobs (environment variables that switch a path off for A/B timing or print statistics) exist only in
// builds made with -DGCANET_MEASUREMENT_AIDS (python -m gcanet_b200.build --aids); the shipped library never reads the
// environment, so no call can be steered into a debug mode that leaves its outputs unwritten.
#ifdef GCANET_MEASUREMENT_AIDS
#include <stdlib.h>
#define GCANET_AID_ENV(name) getenv(name)
#else
#define GCANET_AID_ENV(name) (static_cast<const char *>(nullptr))
#endif

constexpr int kNumSMs = 148;          // B200
constexpr size_t kAlign = 256;        // every workspace sub-buffer starts on this boundary

__host__ __device__ inline size_t align_up(size_t v, size_t a = kAlign) { return (v + a - 1) / a * a; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Bump allocator over a caller-provided workspace (base may be null when only sizing).
struct Carver {
    uintptr_t base;
    size_t off = 0;
    explicit Carver(void *p) : base(reinterpret_cast<uintptr_t>(p)) {}
    template <typename T>
    T *take(size_t count) {
        T *p = reinterpret_cast<T *>(base + off);
        off += align_up(count * sizeof(T));
        return p;
    }
};

inline cudaStream_t as_stream(gcanet_stream_t s) { return static_cast<cudaStream_t>(s); }

// layout.cu
int launch_cn_to_nc(const float *x_cn, float *x_nc, int B, int C, int N, int ld, cudaStream_t st, const float *add_nc = nullptr);
int launch_nc_to_cn(const float *x_nc, float *x_cn, int B, int C, int N, int ld, cudaStream_t st);

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

// Position of the cell (x0, x1, x2), `bits` <= 10 bits per axis, along a 3-D Hilbert curve (Skilling's transform, then the
// usual bit interleave).  The spatial sorts of the kNN kernels use it instead of a Morton code: consecutive cells of a
// Hilbert curve are always adjacent in space, so a run of 64 sorted points never straddles one of the Morton curve's
// jumps and its bounding box stays small -- a query tile then finds ~20 % fewer key tiles within reach.  The order only
// shapes the tiles; exactness rests on their boxes.
__device__ __forceinline__ unsigned hilbert3(unsigned x0, unsigned x1, unsigned x2, int bits) {
    unsigned X[3] = {x0, x1, x2};
    const unsigned M = 1u << (bits - 1);
    for (unsigned Q = M; Q > 1; Q >>= 1) {
        const unsigned P = Q - 1;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (X[i] & Q) X[0] ^= P;
            else { const unsigned t = (X[0] ^ X[i]) & P; X[0] ^= t; X[i] ^= t; }
        }
    }
    X[1] ^= X[0];
    X[2] ^= X[1];
    unsigned t = 0;
    for (unsigned Q = M; Q > 1; Q >>= 1)
        if (X[2] & Q) t ^= Q - 1;
    unsigned code = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        unsigned v = (X[i] ^ t) & 0x3ffu;                   // 10 bits -> every third bit
        v = (v | (v << 16)) & 0x030000ffu;
        v = (v | (v << 8)) & 0x0300f00fu;
        v = (v | (v << 4)) & 0x030c30c3u;
        v = (v | (v << 2)) & 0x09249249u;
        code |= v << (2 - i);
    }
    return code;
}

}  // namespace gcanet
