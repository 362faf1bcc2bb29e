// Shared host/device helpers for the fusion_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cmath>

#include "../../include/fusion_b200.h"

namespace fz {

// ---------------------------------------------------------------------------------------------
// error plumbing: every C-ABI entry point returns 0 / negative and leaves a message here
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);

#define FZ_REQUIRE(cond, ...)                         \
    do {                                              \
        if (!(cond)) {                                \
            ::fz::set_error(__VA_ARGS__);             \
            return FZ_ERR_ARG;                        \
        }                                             \
    } while (0)

#define FZ_CUDA(expr)                                                                   \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            ::fz::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                            __FILE__, __LINE__);                                        \
            return FZ_ERR_CUDA;                                                         \
        }                                                                               \
    } while (0)

#define FZ_LAUNCH_CHECK() FZ_CUDA(cudaGetLastError())

// Optional per-kernel timing (fz_profile_enable): CUDA events recorded on the launching stream around each
// instrumented launch; fz_profile_summary synchronises them and reports count / total ms per kernel name.
void prof_begin(const char* name, cudaStream_t stream);
void prof_end(cudaStream_t stream);
struct ProfScope {
    cudaStream_t s;
    ProfScope(const char* name, cudaStream_t stream) : s(stream) { prof_begin(name, stream); }
    ~ProfScope() { prof_end(s); }
};

// Per-role cycle counters of the persistent tensor-core kernels (fz_debug_set_stats) are compiled in only with
// -DFZ_KERNEL_STATS: the clock reads sit on the kernels' critical paths.
#ifdef FZ_KERNEL_STATS
#define FZ_CLOCK() clock64()
#else
#define FZ_CLOCK() 0ll
#endif

inline int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

template <typename T>
inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------
// order-preserving float <-> unsigned maps (larger float => larger unsigned)
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t ord32(float f) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float unord32(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}
__host__ __device__ __forceinline__ uint64_t ord64(double d) {
#ifdef __CUDA_ARCH__
    uint64_t u = (uint64_t)__double_as_longlong(d);
#else
    uint64_t u;
    memcpy(&u, &d, 8);
#endif
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double unord64(uint64_t k) {
    uint64_t u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double d;
    memcpy(&d, &u, 8);
    return d;
#endif
}

// A sortable record: descending by `skey`, then descending by `tie`; `payload` rides along.
// Scores map through ord32/ord64 (fp32 keys sit in the low 32 bits... of a 64-bit container);
// "lower doc id first" is tie = ~id, "first insertion first" is tie = ~order.
struct __align__(16) Entry {
    uint64_t skey;
    uint32_t tie;
    uint32_t payload;
};
__host__ __device__ __forceinline__ bool entry_before(const Entry& a, const Entry& b) {
    return a.skey > b.skey || (a.skey == b.skey && a.tie > b.tie);
}

#ifdef __CUDACC__
// In-place bitonic sort, "best first", of n_pow2 entries reachable through a generic pointer
// (shared or global memory), by all threads of one CTA.
__device__ __forceinline__ void bitonic_sort_cta(Entry* a, int n_pow2) {
    for (int k = 2; k <= n_pow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
                int p = i ^ j;
                if (p > i) {
                    Entry x = a[i], y = a[p];
                    bool up = (i & k) == 0;            // this run is sorted "best first"
                    if (entry_before(y, x) == up) {
                        a[i] = y;
                        a[p] = x;
                    }
                }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif

}  // namespace fz
