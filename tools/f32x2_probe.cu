// Stand-alone probe: issue rate of the packed FP32 instructions of sm_100a (FFMA2 / FADD2 / FMUL2,
// PTX fma.rn.f32x2 ...) against scalar FFMA, per SM sub-partition.  Decides whether the FFT /
// contraction code should be written on (re, im) register pairs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_probe tools/f32x2_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

typedef unsigned long long u64;

template <int MODE>
__global__ void __launch_bounds__(256) probe(float *out, int iters, long long *cycles) {
    constexpr int NCH = 8;  // independent chains per thread
    float a[2 * NCH];
    for (int i = 0; i < 2 * NCH; ++i) a[i] = threadIdx.x * 1e-3f + i;
    const float b = 1.0000001f, c = 1e-9f;
    u64 p[NCH];
    for (int i = 0; i < NCH; ++i) asm("mov.b64 %0, {%1,%2};" : "=l"(p[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
    u64 pb, pc;
    asm("mov.b64 %0, {%1,%1};" : "=l"(pb) : "f"(b));
    asm("mov.b64 %0, {%1,%1};" : "=l"(pc) : "f"(c));
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {  // 16 scalar FFMA
#pragma unroll
            for (int i = 0; i < 2 * NCH; ++i) a[i] = fmaf(a[i], b, c);
        } else if (MODE == 1) {  // 8 FFMA2 (same flops as mode 0)
#pragma unroll
            for (int i = 0; i < NCH; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pb), "l"(pc));
        } else if (MODE == 2) {  // 8 FADD2
#pragma unroll
            for (int i = 0; i < NCH; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pc));
        } else if (MODE == 3) {  // 16 scalar FADD
#pragma unroll
            for (int i = 0; i < 2 * NCH; ++i) a[i] = a[i] + c;
        } else if (MODE == 4) {  // 8 FFMA2 with swapped / negated operand (complex multiply shape)
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
                float lo, hi;
                asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i]));
                u64 sw;
                asm("mov.b64 %0, {%1,%2};" : "=l"(sw) : "f"(-hi), "f"(lo));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p[i]) : "l"(sw), "l"(pb), "l"(pc));
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < 2 * NCH; ++i) s += a[i];
    for (int i = 0; i < NCH; ++i) {
        float lo, hi;
        asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i]));
        s += lo + hi;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int MODE>
static void run(const char *name, int instr_per_iter, int flops_per_instr) {
    float *out;
    long long *cyc, h;
    cudaMalloc(&out, 148 * 4 * 256 * sizeof(float));  // 4 CTAs per SM at most
    cudaMalloc(&cyc, sizeof(long long));
    const int iters = 20000;
    for (int cps = 1; cps <= 4; cps *= 2) {  // 256-thread CTAs per SM: 2 warps per sub-partition each
        const int ctas = 148 * cps;
        probe<MODE><<<ctas, 256>>>(out, 100, cyc);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0);
        probe<MODE><<<ctas, 256>>>(out, iters, cyc);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        const double warp_instr = (double)iters * instr_per_iter * (cps * 2);  // per sub-partition
        printf("%-28s warps/SMSP=%d  %.3f warp-instr/clk/SMSP  %.1f TFLOP/s (whole GPU, %.3f ms)\n", name,
               cps * 2, warp_instr / (double)h,
               (double)ctas * 256 * iters * instr_per_iter * flops_per_instr / (ms * 1e-3) * 1e-12, ms);
    }
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    run<0>("scalar FFMA", 16, 2);
    run<1>("FFMA2", 8, 4);
    run<3>("scalar FADD", 16, 1);
    run<2>("FADD2", 8, 2);
    run<4>("FFMA2 swap+neg operand", 8, 4);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("error %s\n", cudaGetErrorString(e));
        return 1;
    }
    return 0;
}
