// K-A on the 5th-generation tensor cores (tcgen05, accumulator in TMEM) for wide arrays / high SH
// orders, where the analysis GEMM  z[j][t] = sum_c Rm[j][c] x[c][t]  is FP32-pipe bound on SIMT
// (order 8, 110 channels: 23 flop/B).  Included by sh_kernels.cuh.
//
//   D[t][j] (TMEM, fp32)  =  A[t][c] (smem, K-major)  x  B[j][c]^T (smem, K-major)
//   M = 128 samples of one stream, N = rows padded to 16, K = channels padded to 8, kind::tf32.
//
// Plain TF32 (10-bit mantissa) would break the 1e-5 parity bar, so both operands are split
// a = a_hi + a_lo (hi = a rounded to TF32, lo = a - hi, exact in fp32) and three MMAs accumulate
// hi*hi + lo*hi + hi*lo in the fp32 accumulator; measured max error 2e-6 relative
// (tools/tc_probe.cu).  The analysis matrix is split and laid out on the host; the input tile is
// split on the fly while it is scattered from coalesced global loads into the un-swizzled
// core-matrix layout [c/4][t/8][t%8][c%4] (a thread's four channels form one 16-byte run, a warp
// writes 512 contiguous bytes).  MN-major operands would avoid that transposition but produced
// zeros with un-swizzled TF32 descriptors on this part (probe variant 0), so K-major it is.
//
// First version: the three phases of a tile (load+split, MMA, epilogue) are not overlapped; the
// TMA-staged, warp-specialised pipeline is the follow-up.
#pragma once

namespace rtb {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor: no swizzle (layout_type 0), descriptor version 1 (Blackwell);
// offsets are encoded without their 4 LSBs
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
}

__device__ __forceinline__ void split_tf32(float v, float &hi, float &lo) {
    hi = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);  // round to nearest TF32
    lo = v - hi;
}

#define RTB_TC_M 128
#define RTB_TC_MAX_K 128  // channels (padded) the register prefetch is sized for

__global__ void __launch_bounds__(256, 1)
sh_analysis_tc_kernel(const float *__restrict__ x,     // [S][C][B]
                      const float *__restrict__ bsrc,  // [2][Kpad/4][Npad/8][8][4]: hi then lo
                      float *__restrict__ zcur,        // [S][J][B]
                      int S, int C, int J, int B, int Kpad, int Npad, int tmem_cols) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *a_hi = reinterpret_cast<float *>(smem_raw);   // [Kpad/4][16][8][4]
    float *a_lo = a_hi + (size_t)Kpad * RTB_TC_M;
    float *b_hi = a_lo + (size_t)Kpad * RTB_TC_M;        // [Kpad/4][Npad/8][8][4]
    float *b_lo = b_hi + (size_t)Kpad * Npad;
    __shared__ __align__(8) uint64_t mma_bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < Kpad * Npad / 2; i += 256)  // both halves (hi, lo) are contiguous
        reinterpret_cast<float4 *>(b_hi)[i] = __ldg(reinterpret_cast<const float4 *>(bsrc) + i);
    if (tid == 0) {
        mbar_init(&mma_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    // instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3, M >> 4
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(Npad >> 3) << 17) |
                           ((uint32_t)(RTB_TC_M >> 4) << 24);
    const uint32_t lbo_a = (RTB_TC_M / 8) * 128, lbo_b = (Npad / 8) * 128, sbo = 128;
    const int tiles_per_stream = B / RTB_TC_M;
    const long long n_tiles = (long long)S * tiles_per_stream;
    const int t_in = tid & 127, khalf = tid >> 7;  // loader: sample, and which half of the k-chunks
    unsigned parity = 0;

    // The tile of the NEXT iteration is fetched into registers right after the MMAs of the current
    // one are issued, so its HBM latency hides behind the tensor-core work and the epilogue.
    constexpr int NKC = RTB_TC_MAX_K / 8;  // k-chunks (of 4 channels) per loader thread, at most
    float xr[NKC][4];
    auto load_tile = [&](long long tile) {
        const long long s = tile / tiles_per_stream;
        const int t0 = (int)(tile % tiles_per_stream) * RTB_TC_M;
        const float *xs = x + ((size_t)s * C) * B + t0 + t_in;
#pragma unroll
        for (int u = 0; u < NKC; ++u) {
            const int kc = khalf + 2 * u;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = kc * 4 + i;
                xr[u][i] = (kc < Kpad / 4 && c < C) ? __ldg(xs + (size_t)c * B) : 0.f;
            }
        }
    };
    if (blockIdx.x < n_tiles) load_tile(blockIdx.x);

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long s = tile / tiles_per_stream;
        const int t0 = (int)(tile % tiles_per_stream) * RTB_TC_M;
        // ---- phase 1: split the prefetched samples, scatter into the core-matrix layout
#pragma unroll
        for (int u = 0; u < NKC; ++u) {
            const int kc = khalf + 2 * u;
            if (kc < Kpad / 4) {
                float hi[4], lo[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) split_tf32(xr[u][i], hi[i], lo[i]);
                const size_t off = (((size_t)kc * (RTB_TC_M / 8) + (t_in >> 3)) * 8 + (t_in & 7)) * 4;
                *reinterpret_cast<float4 *>(a_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4 *>(a_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> tensor core reads
        __syncthreads();
        // ---- phase 2: one thread issues hi*hi + lo*hi + hi*lo over all k-steps
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t accumulate = 0;
            for (int pass = 0; pass < 3; ++pass) {
                const uint32_t pa = smem_u32(pass == 1 ? a_lo : a_hi);
                const uint32_t pb = smem_u32(pass == 2 ? b_lo : b_hi);
                for (int ks = 0; ks < Kpad / 8; ++ks) {
                    const uint64_t da = umma_desc(pa + ks * 2 * lbo_a, lbo_a, sbo);
                    const uint64_t db = umma_desc(pb + ks * 2 * lbo_b, lbo_b, sbo);
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                        ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
                    accumulate = 1;
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                         ::"r"(smem_u32(&mma_bar)) : "memory");
        }
        if (tile + gridDim.x < n_tiles) load_tile(tile + gridDim.x);
        // ---- phase 3: TMEM -> registers -> z rows.  Warp w reads the lane quarter w % 4 (rows t);
        // the two warps of a quarter alternate over 32-column chunks
        mbar_wait(&mma_bar, parity);
        parity ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;
        float *zt = zcur + ((size_t)s * J) * B + t0 + q * 32 + lane;
        for (int cb = (warp >> 2) * 32; cb < Npad; cb += 64) {
            uint32_t v[32];
            const uint32_t addr = tmem + ((uint32_t)(q * 32) << 16) + cb;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(addr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (cb + j < J) zt[(size_t)(cb + j) * B] = __uint_as_float(v[j]);
        }
        // the accumulator and the A buffers are reused by the next tile
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols));
}

}  // namespace rtb
