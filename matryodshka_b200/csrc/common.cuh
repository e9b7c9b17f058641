// Shared host/device helpers for the msi_b200 kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/msi_b200.h"

namespace msi {

// thread-local last-error text (msi_last_error)
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

#define MSI_CHECK_ARG(cond, ...)              \
    do {                                      \
        if (!(cond)) {                        \
            msi::set_error(__VA_ARGS__);      \
            return MSI_ERR_INVALID_ARG;       \
        }                                     \
    } while (0)

#define MSI_CUDA(expr)                                                                  \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            msi::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                           __FILE__, __LINE__);                                         \
            return MSI_ERR_CUDA;                                                        \
        }                                                                               \
    } while (0)

#define MSI_LAUNCH_CHECK()                                                              \
    do {                                                                                \
        cudaError_t _e = cudaGetLastError();                                            \
        if (_e != cudaSuccess) {                                                        \
            msi::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),  \
                           __FILE__, __LINE__);                                         \
            return MSI_ERR_CUDA;                                                        \
        }                                                                               \
        msi::count_launch();                                                            \
    } while (0)

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// fp16 hi/lo split of a (scaled) float: hi = rn(x), lo = rn(x - hi).  ~22 significand bits.
__device__ __forceinline__ void split_half(float x, __half& hi, __half& lo) {
    hi = __float2half_rn(x);
    lo = __float2half_rn(x - __half2float(hi));
}

}  // namespace msi
