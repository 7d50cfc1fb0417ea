// Probe: tcgen05 kind::tf32 with an MN-major A operand in the 128-byte-swizzled canonical layout,
// i.e. the layout a TMA tile load {32 samples x channels} with SWIZZLE_128B leaves in shared memory.
//   D[t][j] = sum_c x[c][t] * R[j][c],  M = 128 (t), N = 32 (j), K = 32 (c)
// A: sample block mb (32 samples), channel c, sample s:  byte = mb*BLK + (c/8)*1024 + (c%8)*128
//    + (((s/4) ^ (c%8)) * 16) + (s%4)*4   (Swizzle<3,4,3>), BLK = (K/8)*1024.
// B: K-major, un-swizzled core matrices [c/4][j/8][j%8][c%4] (as in sh_analysis_tc.cuh).
// Variants sweep the LBO / SBO assignment of the A descriptor.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/tc_probe_mn tools/tc_probe_mn.cu
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

constexpr int M = 128, N = 32, K = 32;
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout,
                                              uint32_t lbo_mode = 0) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46) | ((uint64_t)(lbo_mode & 1) << 52) |
           ((uint64_t)layout << 61);
}

__global__ void __launch_bounds__(128) probe(const float *x, const float *bsrc, float *d_out, int variant) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float *a = reinterpret_cast<float *>(smem);                 // 16 KB, swizzled MN-major
    float *b = a + K * M;                                       // [K/4][N/8][8][4]
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr uint32_t BLK = (K / 8) * 1024;
    for (int i = tid; i < K * M; i += 128) {
        const int c = i / M, t = i % M, mb = t / 32, s = t % 32;
        uint32_t off = mb * BLK + (c / 8) * 1024 + (c % 8) * 128 + (((s / 4) ^ (c % 8)) * 16) + (s % 4) * 4;
        const int lay_variant = variant == 7 ? 5 : (variant == 8 ? 0 : variant);
        if (lay_variant >= 5)  // un-swizzled MN-major core matrices: 4 samples x 8 channels, 128 B each
            off = (c / 8) * (M / 4) * 128 + (t / 4) * 128 + (c % 8) * 16 + (t % 4) * 4;
        a[off / 4] = x[i];
    }
    for (int i = tid; i < K * N; i += 128) b[i] = bsrc[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(32));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        const uint32_t lbo_mode = variant >= 7 ? 1u : 0u;  // round 2: same layouts with the descriptor's lbo_mode bit
        if (variant == 7) variant = 5;
        if (variant == 8) variant = 0;
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (variant == 4 ? 0u : (1u << 15)) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        // A: per MMA (K = 8 channels) the start address advances by 1024 B (one 8-channel row group)
        uint32_t lbo_a = BLK, sbo_a = 1024;
        if (variant == 1) { lbo_a = 1024; sbo_a = BLK; }
        if (variant == 2) { lbo_a = BLK; sbo_a = 128; }
        if (variant == 3) { lbo_a = 128; sbo_a = BLK; }
        const uint32_t lbo_b = (N / 8) * 128, sbo_b = 128;
        uint32_t kstep = 1024, layout = 2 /* SWIZZLE_128B */;
        if (variant == 4) { layout = 0; lbo_a = (M / 8) * 128; sbo_a = 128; }
        if (variant >= 5) {  // interleaved MN-major: sample groups 128 B apart, 8-channel groups (M/4)*128 B apart
            layout = 0; kstep = (M / 4) * 128;
            lbo_a = (variant == 5) ? kstep : 128;
            sbo_a = (variant == 5) ? 128 : kstep;
        }
        for (int s = 0; s < K / 8; ++s) {
            const uint64_t da = make_desc(smem_u32(a) + s * kstep, lbo_a, sbo_a, layout, lbo_mode);
            const uint64_t db = make_desc(smem_u32(b) + s * 2 * lbo_b, lbo_b, sbo_b, 0);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(s ? 1u : 0u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tWAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE;\n\tbra WAIT;\n\tDONE:\n\t}"
                 ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[32];
    const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) d_out[(size_t)tid * N + j] = __uint_as_float(v[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32));
}

int main() {
    std::vector<float> x((size_t)K * M), r((size_t)N * K), bsrc((size_t)K * N);
    srand(1);
    auto tf = [](float v) { uint32_t u; memcpy(&u, &v, 4); u &= 0xffffe000u; float o; memcpy(&o, &u, 4); return o; };
    for (auto &v : x) v = tf(rand() / (float)RAND_MAX - 0.5f);   // exactly representable in TF32
    for (auto &v : r) v = tf(rand() / (float)RAND_MAX - 0.5f);
    for (int j = 0; j < N; ++j)
        for (int c = 0; c < K; ++c)
            bsrc[(((size_t)(c / 4) * (N / 8) + j / 8) * 8 + j % 8) * 4 + c % 4] = r[(size_t)j * K + c];
    float *dx, *db, *dd;
    cudaMalloc(&dx, x.size() * 4); cudaMalloc(&db, bsrc.size() * 4); cudaMalloc(&dd, (size_t)M * N * 4);
    cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(db, bsrc.data(), bsrc.size() * 4, cudaMemcpyHostToDevice);
    const int smem = (K * M + K * N) * 4 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<float> d((size_t)M * N);
    for (int variant = 0; variant < 9; ++variant) {
        cudaMemset(dd, 0, (size_t)M * N * 4);
        probe<<<1, 128, smem>>>(dx, db, dd, variant);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant %d CUDA error: %s\n", variant, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
        double worst = 0, peak = 0;
        int bad_rows = 0;
        for (int t = 0; t < M; ++t) {
            double rw = 0;
            for (int j = 0; j < N; ++j) {
                double ref = 0;
                for (int c = 0; c < K; ++c) ref += (double)x[(size_t)c * M + t] * r[(size_t)j * K + c];
                rw = fmax(rw, fabs(ref - d[(size_t)t * N + j]));
                peak = fmax(peak, fabs(ref));
            }
            worst = fmax(worst, rw);
            bad_rows += rw > 1e-4;
        }
        printf("variant %d: max abs err %.3e (peak %.3f), rows off %d/%d, d[0]=%f d[N]=%f\n", variant, worst, peak, bad_rows, M, d[0], d[N]);
    }
    return 0;
}
