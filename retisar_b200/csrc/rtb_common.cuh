// Shared host/device helpers of libretisar_b200 (error reporting, complex arithmetic).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstdint>

#include "../../include/retisar_b200.h"

namespace rtb {

// thread-local message of the last failure, returned by rtb_last_error()
char *last_error_buffer();
int fail(int code, const char *fmt, ...);

#define RTB_CUDA(expr)                                                                      \
    do {                                                                                    \
        cudaError_t err__ = (expr);                                                         \
        if (err__ != cudaSuccess)                                                           \
            return rtb::fail(RTB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                  \
                             cudaGetErrorString(err__), __FILE__, __LINE__);                \
    } while (0)

// RAII device switch so a plan can be driven from a thread whose current device differs
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

__host__ __device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// acc += a * b
__device__ __forceinline__ void cfma(float2 &acc, float2 a, float2 b) {
    acc.x = fmaf(a.x, b.x, acc.x);
    acc.x = fmaf(-a.y, b.y, acc.x);
    acc.y = fmaf(a.x, b.y, acc.y);
    acc.y = fmaf(a.y, b.x, acc.y);
}

}  // namespace rtb
