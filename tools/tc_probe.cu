// Stand-alone probe of the tcgen05 (5th-gen tensor core) path used by the analysis GEMM:
//   D[t][j] = sum_c x[c][t] * R[j][c],   M = 128 (t), N = 96 (j), K = 112 (c)
// A = x^T as an MN-major operand, B = R as a K-major operand, both un-swizzled "interleaved"
// core-matrix layouts; kind::tf32 with the 3-way split (hi*hi + lo*hi + hi*lo) for fp32 accuracy.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tools/tc_probe.cu
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int M = 128, N = 96, K = 112;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, no swizzle (layout_type 0), version 1 (Blackwell)
__device__ int g_variant = 0;
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    if (!(g_variant & 1)) d |= (uint64_t)1 << 46;
    return d;
}

__global__ void __launch_bounds__(128) probe(const float *x /*[K][M]*/, const float *bhi, const float *blo,
                                             float *d_out /*[M][N]*/, int split) {
    extern __shared__ __align__(128) unsigned char smem[];
    float *a_hi = reinterpret_cast<float *>(smem);           // [K/8][32][8][4]
    float *a_lo = a_hi + K * M;
    float *b_hi = a_lo + K * M;                              // [K/4][N/8][8][4]
    float *b_lo = b_hi + K * N;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;

    // operand A (K-major, un-swizzled core matrices [c/4][t/8][t%8][c%4]): thread t gathers 4
    // consecutive channels of its sample (coalesced 128-byte row reads) and stores one 16-byte run
    for (int kc = 0; kc < K / 4; ++kc) {
        const int t = tid;
        float v[4], hi[4], lo[4];
        for (int i = 0; i < 4; ++i) v[i] = x[(size_t)(kc * 4 + i) * M + t];
        for (int i = 0; i < 4; ++i) {
            hi[i] = __uint_as_float(__float_as_uint(v[i]) & 0xffffe000u);
            lo[i] = v[i] - hi[i];
        }
        const size_t off = (((size_t)kc * (M / 8) + t / 8) * 8 + t % 8) * 4;
        *reinterpret_cast<float4 *>(a_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4 *>(a_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
    for (int i = tid; i < K * N / 4; i += 128) {
        reinterpret_cast<float4 *>(b_hi)[i] = reinterpret_cast<const float4 *>(bhi)[i];
        reinterpret_cast<float4 *>(b_lo)[i] = reinterpret_cast<const float4 *>(blo)[i];
    }
    if (split == 3) {  // debug: all-ones operands, D must equal K everywhere
        for (int i = tid; i < K * M; i += 128) { a_hi[i] = 1.0f; a_lo[i] = 1.0f; }
        for (int i = tid; i < K * N; i += 128) { b_hi[i] = 1.0f; b_lo[i] = 1.0f; }
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // generic-proxy writes -> visible to the async proxy (tensor core reads shared memory)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (split == 2) {  // debug: TMEM store/load round trip, no MMA
        for (int cb = 0; cb < N; cb += 32) {
            const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + cb;
            for (int j = 0; j < 32; ++j) {
                const uint32_t val = __float_as_uint((float)((warp * 32 + (tid & 31)) * 1000 + cb + j));
                asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(addr + j), "r"(val) : "memory");
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }

    if (tid == 0 && split != 2) {
        // instruction descriptor: D = F32, A = B = TF32, A MN-major, B K-major, N, M
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((g_variant & 2) ? 0u : (1u << 15)) | (0u << 16) |
                               ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t lbo_a = (M / 8) * 128, sbo_a = 128;   // A: k-chunks (M/8)*128 B apart, m-groups 128 B apart
        const uint32_t lbo_b = (N / 8) * 128, sbo_b = 128;   // B: k-chunks (N/8)*128 B apart, n-groups 128 B apart
        const int npass = (split == 1) ? 3 : 1;
        uint32_t accumulate = 0;
        for (int pass = 0; pass < npass; ++pass) {
            const float *pa = (pass == 1) ? a_lo : a_hi;
            const float *pb = (pass == 2) ? b_lo : b_hi;
            for (int s = 0; s < K / 8; ++s) {
                const uint64_t da = make_desc(smem_u32(pa) + s * 2 * lbo_a, lbo_a, sbo_a);
                const uint64_t db = make_desc(smem_u32(pb) + s * 2 * lbo_b, lbo_b, sbo_b);
                if (g_variant & 4) {
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                        ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0) : "memory");
                } else {
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                        ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
                }
                accumulate = 1;
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // wait for the MMAs
    if (split != 2) asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE;\n\tbra WAIT;\n\tDONE:\n\t}"
        ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (g_variant & 8) { for (int i = 0; i < 2000; ++i) __nanosleep(1000); }

    // epilogue: warp w owns TMEM lanes 32w..32w+31 (rows t), 32 columns per load
    for (int cb = 0; cb < N; cb += 32) {
        uint32_t v[32];
        const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + cb;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
            "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(addr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int t = warp * 32 + (tid & 31);
        for (int j = 0; j < 32; ++j) d_out[(size_t)t * N + cb + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

int main() {
    std::vector<float> x((size_t)K * M), r((size_t)N * K), bhi((size_t)K * N), blo((size_t)K * N);
    srand(1);
    for (auto &v : x) v = (rand() / (float)RAND_MAX - 0.5f);
    for (auto &v : r) v = (rand() / (float)RAND_MAX - 0.5f);
    // B in the K-major interleaved layout [K/4][N/8][8][4]
    for (int j = 0; j < N; ++j)
        for (int c = 0; c < K; ++c) {
            const float v = r[(size_t)j * K + c];
            uint32_t u; memcpy(&u, &v, 4); u &= 0xffffe000u;
            float hi; memcpy(&hi, &u, 4);
            const size_t off = (((size_t)(c / 4) * (N / 8) + j / 8) * 8 + j % 8) * 4 + c % 4;
            bhi[off] = hi; blo[off] = v - hi;
        }
    float *dx, *dbhi, *dblo, *dd;
    cudaMalloc(&dx, x.size() * 4); cudaMalloc(&dbhi, bhi.size() * 4); cudaMalloc(&dblo, blo.size() * 4);
    cudaMalloc(&dd, (size_t)M * N * 4);
    cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dbhi, bhi.data(), bhi.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dblo, blo.data(), blo.size() * 4, cudaMemcpyHostToDevice);
    const int smem = (2 * K * M + 2 * K * N) * 4;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<float> d((size_t)M * N);
    for (int variant = 2; variant < 3; ++variant)
    for (int split = 3; split >= 0; --split) {
        cudaMemcpyToSymbol(g_variant, &variant, sizeof(int));
        printf("variant %d: ", variant);
        cudaMemset(dd, 0, (size_t)M * N * 4);
        probe<<<1, 128, smem>>>(dx, dbhi, dblo, dd, split);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
        double worst = 0, peak = 0;
        for (int t = 0; t < M; ++t)
            for (int j = 0; j < N; ++j) {
                double ref = 0;
                for (int c = 0; c < K; ++c) ref += (double)x[(size_t)c * M + t] * r[(size_t)j * K + c];
                worst = fmax(worst, fabs(ref - d[(size_t)t * N + j]));
                peak = fmax(peak, fabs(ref));
            }
        int nz = 0; for (float v : d) nz += (v != 0.f);
        printf("split=%d max abs err %.3e (peak %.3f) nonzeros %d/%d d[0]=%f d[1]=%f d[N]=%f\n", split, worst, peak, nz, (int)d.size(), d[0], d[1], d[N]);
    }
    return 0;
}
