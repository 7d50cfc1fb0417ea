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

// Function attributes are per device: set them once per (kernel, device) pair, so that plans on
// several devices of one process all get their large dynamic shared-memory windows.
bool first_use_on_device(const void *func);  // capi.cu
template <class K>
inline void kernel_attrs_once(K kernel, int max_dynamic_smem, bool prefer_smem_carveout = false) {
    if (!first_use_on_device(reinterpret_cast<const void *>(kernel))) return;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dynamic_smem);
    if (prefer_smem_carveout) cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
}

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization
// attribute may start while its predecessor in the stream is still draining; everything before
// pdl_wait() (table staging, barrier setup) overlaps the predecessor's tail, everything after it sees
// the predecessor's memory.  Both are no-ops for a normally launched kernel.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              bool pdl, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Packed FP32 (sm_100a FFMA2 / FADD2 / FMUL2, PTX *.f32x2): a complex value lives in an aligned
// register pair and every complex add / scale / multiply-accumulate is one or two instructions.
// The FP32 pipe does the same flops per cycle either way (tools/f32x2_probe.cu: 70.6 vs 71.3
// TFLOP/s), but the packed forms need half the issue slots -- and the FFT / contraction kernels
// are issue bound.  ptxas folds the (re, im) swaps, negations and scalar broadcasts below into
// operand modifiers (.F32x2.LO_HI, .NP, .F32), so they cost nothing.
#ifndef RTB_PACKED_F32
#define RTB_PACKED_F32 1
#endif
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pk2(float lo, float hi) {
    f32x2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f32x2_t pk2(float2 a) { return pk2(a.x, a.y); }
__device__ __forceinline__ float2 up2(f32x2_t v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ f32x2_t add2(f32x2_t a, f32x2_t b) {
    f32x2_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t mul2(f32x2_t a, f32x2_t b) {
    f32x2_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c) {
    f32x2_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

#if RTB_PACKED_F32
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return up2(add2(pk2(a), pk2(b))); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return up2(add2(pk2(a), pk2(-b.x, -b.y))); }
// a * s (real scalar)
__device__ __forceinline__ float2 cscale(float2 a, float s) { return up2(mul2(pk2(a), pk2(s, s))); }
// a * b = b * (a.x, a.x) + (-b.y, b.x) * (a.y, a.y)
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return up2(fma2(pk2(-b.y, b.x), pk2(a.y, a.y), mul2(pk2(b), pk2(a.x, a.x))));
}
// acc += a * b
__device__ __forceinline__ void cfma(float2 &acc, float2 a, float2 b) {
    acc = up2(fma2(pk2(-b.y, b.x), pk2(a.y, a.y), fma2(pk2(b), pk2(a.x, a.x), pk2(acc))));
}
#else
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// acc += a * b
__device__ __forceinline__ void cfma(float2 &acc, float2 a, float2 b) {
    acc.x = fmaf(a.x, b.x, acc.x);
    acc.x = fmaf(-a.y, b.y, acc.x);
    acc.y = fmaf(a.x, b.y, acc.y);
    acc.y = fmaf(a.y, b.x, acc.y);
}
#endif

}  // namespace rtb
