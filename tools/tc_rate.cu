// tcgen05 kind::tf32 issue-rate and layout probe for the analysis GEMM (M = 128, K-major operands):
//   layout 0: un-swizzled core matrices [k/4][row/8][8][4]        (what sh_analysis_tc_kernel uses)
//   layout 1: 128-byte swizzle, atoms of 8 rows x 128 B (32 tf32), 16-byte chunks XORed with row % 8
// For every layout and N in {32, 96, 176, 256}: checks D = A * B^T against the host (single TF32
// pass on TF32-rounded inputs, so the only error is fp32 accumulation order) and times a chain of
// back-to-back MMAs (clock64 from the first issue to the commit's arrival) -> cycles per MMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/tc_rate tools/tc_rate.cu
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

constexpr int M = 128, K = 64, NMAX = 256;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46) | ((uint64_t)layout_type << 61);
}

// element offset (floats) of (row, k) in an operand with `rows` rows
__host__ __device__ inline size_t op_off(int layout, int rows, int row, int k) {
    if (layout == 0) return (((size_t)(k / 4) * (rows / 8) + row / 8) * 8 + row % 8) * 4 + k % 4;
    const int atom_col = k / 32, kk = k % 32;
    return (size_t)atom_col * rows * 32 + (size_t)(row / 8) * 256 + (row % 8) * 32 + (((kk / 4) ^ (row % 8)) * 4) + kk % 4;
}

__global__ void __launch_bounds__(128) probe(const float *a_src, const float *b_src, float *d_out, long long *cycles,
                                             int layout, int N, int reps) {
    extern __shared__ __align__(1024) unsigned char smem[];
    // 128 x 64 floats = 32 KB, on a 1024-byte boundary of the shared window (swizzle atoms)
    float *a_s = reinterpret_cast<float *>(smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u));
    float *b_s = a_s + M * K;                      // N x 64
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < M * K; i += 128) a_s[i] = a_src[i];
    for (int i = tid; i < N * K; i += 128) b_s[i] = b_src[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r)
            for (int s = 0; s < K / 8; ++s) {
                uint64_t da, db;
                if (layout == 0) {
                    const uint32_t lbo_a = (M / 8) * 128, lbo_b = (N / 8) * 128;
                    da = make_desc(smem_u32(a_s) + s * 2 * lbo_a, lbo_a, 128, 0);
                    db = make_desc(smem_u32(b_s) + s * 2 * lbo_b, lbo_b, 128, 0);
                } else {  // 4 k-steps of 32 B inside a 128-byte swizzle atom, then the next atom column
                    da = make_desc(smem_u32(a_s) + (s / 4) * M * 128 + (s % 4) * 32, 16, 1024, 2);
                    db = make_desc(smem_u32(b_s) + (s / 4) * N * 128 + (s % 4) * 32, 16, 1024, 2);
                }
                const uint32_t acc = (r | s) ? 1u : 0u;
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                    ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
            }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        long long t1 = clock64();
        asm volatile(
            "{\n\t.reg .pred p;\n\tWAIT0:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE0;\n\tbra WAIT0;\n\tDONE0:\n\t}"
            ::"r"(smem_u32(&bar)), "r"(0) : "memory");
        long long t2 = clock64();
        cycles[0] = t1 - t0;
        cycles[1] = t2 - t0;
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int cb = 0; cb < N; cb += 16) {
        uint32_t v[16];
        const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + cb;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(addr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int t = warp * 32 + (tid & 31);
        for (int j = 0; j < 16; ++j) d_out[(size_t)t * N + cb + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

static float tf32_round(float v) {
    uint32_t u; memcpy(&u, &v, 4); u = (u + 0x1000u) & 0xffffe000u; memcpy(&v, &u, 4); return v;
}

int main() {
    std::vector<float> a((size_t)M * K), b((size_t)NMAX * K);
    srand(1);
    for (auto &v : a) v = tf32_round(rand() / (float)RAND_MAX - 0.5f);
    for (auto &v : b) v = tf32_round(rand() / (float)RAND_MAX - 0.5f);
    float *da, *db, *dd; long long *dc;
    cudaMalloc(&da, a.size() * 4); cudaMalloc(&db, b.size() * 4); cudaMalloc(&dd, (size_t)M * NMAX * 4); cudaMalloc(&dc, 16);
    const int smem = (M * K + NMAX * K) * 4 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int ns[4] = {32, 96, 176, 256};
    for (int layout = 0; layout < 2; ++layout)
        for (int ni = 0; ni < 4; ++ni) {
            const int N = ns[ni];
            std::vector<float> al((size_t)M * K, 0.f), bl((size_t)N * K, 0.f), d((size_t)M * N);
            for (int r = 0; r < M; ++r) for (int k = 0; k < K; ++k) al[op_off(layout, M, r, k)] = a[(size_t)r * K + k];
            for (int r = 0; r < N; ++r) for (int k = 0; k < K; ++k) bl[op_off(layout, N, r, k)] = b[(size_t)r * K + k];
            cudaMemcpy(da, al.data(), al.size() * 4, cudaMemcpyHostToDevice);
            cudaMemcpy(db, bl.data(), bl.size() * 4, cudaMemcpyHostToDevice);
            for (int reps : {1, 32}) {
                cudaMemset(dd, 0, (size_t)M * N * 4);
                probe<<<1, 128, smem>>>(da, db, dd, dc, layout, N, reps);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
                long long cyc[2];
                cudaMemcpy(cyc, dc, 16, cudaMemcpyDeviceToHost);
                cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
                double worst = 0;
                for (int t = 0; t < M; ++t)
                    for (int j = 0; j < N; ++j) {
                        double ref = 0;
                        for (int c = 0; c < K; ++c) ref += (double)a[(size_t)t * K + c] * b[(size_t)j * K + c];
                        worst = fmax(worst, fabs(ref * reps - d[(size_t)t * N + j]));
                    }
                const int n_mma = reps * K / 8;
                printf("layout %d N %3d reps %2d: max abs err %.3e | issue %.1f cyc/MMA, complete %.1f cyc/MMA (%d MMAs)\n",
                       layout, N, reps, worst, (double)cyc[0] / n_mma, (double)cyc[1] / n_mma, n_mma);
            }
        }
    return 0;
}
