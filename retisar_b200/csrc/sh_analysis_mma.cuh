// K-A for narrow arrays (C <= 32 channels, e.g. the 32-channel Eigenmike of the headline config) on
// the warp-level tensor-core path: mma.sync.m16n8k8 TF32 with the 3-way split (hi*hi + lo*hi +
// hi*lo, fp32 accumulate) that keeps the 1e-5 parity bar.  Included by sh_kernels.cuh.
//
// Why not tcgen05 here: with K = 32 the UMMA pipeline of sh_analysis_tc.cuh spends more on splitting
// and scattering the operand into the core-matrix layout than the SIMT GEMM spends on FMAs (measured
// 170 vs 141 us).  mma.sync takes the A fragments straight from the [channel][sample] tile in shared
// memory (4 LDS.32 + 8 ALU per 16x8 fragment) and keeps the whole split analysis matrix (32 x 32) in
// registers as B fragments.
//
// Measured on B200 (c2): 132 vs 142 us per block of 10 240 streams, 24.0 vs 24.9 us at 1 024 -- a
// small gain only: the legacy HMMA.1688.F32.TF32 path issues at about the FP32-SIMT flop rate on
// this part (one CTA per SM: 3.5 us per 32-channel tile = 192 HMMA per warp), so with the 3-way
// split the tensor instruction stream is no shorter than the FFMA2 one; what is saved is the
// shared-memory operand traffic of the SIMT micro-tile.  RTB_ANALYSIS_MMA=0 selects the SIMT kernel.
//
//   D[sample][row] = sum_c A[sample][c] * B[c][row],  A = x^T (row-major fragments), B = R^T
//
// Same persistent bulk-copy pipeline as sh_analysis_kernel (full / empty mbarriers), except that the
// tile rows are copied one by one into rows padded by 8 floats: fragment loads a0..a3 of a warp then
// hit 32 different banks.
#pragma once

namespace rtb {

#define RTB_MMA_CP 32        // channels (K), zero padded
#define RTB_MMA_JP 32        // analysis rows per pass (N), zero padded
#define RTB_MMA_TW 256       // samples per tile at most (= block length when B <= 256)
#define RTB_MMA_PAD 8

__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void split_tf32_rna(float v, uint32_t &hi, uint32_t &lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(v));
    lo = __float_as_uint(v - __uint_as_float(hi));  // the MMA reads the upper 19 bits of lo
}

static __global__ void __launch_bounds__(128, 3)
sh_analysis_mma_kernel(const float *__restrict__ x,     // [S][C][B]
                       const float *__restrict__ rsp,   // [n_jchunks][2: hi, lo][JP rows][CP channels]
                       float *__restrict__ zcur,        // [S][J][B]
                       int S, int C, int J, int B, int TW, int n_jchunks, int stages) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[RTB_AN_MAX_STAGES], empty_bar[RTB_AN_MAX_STAGES];
    float *tiles = reinterpret_cast<float *>(smem_raw);
    const int TWP = TW + RTB_MMA_PAD;
    const int tile_elems = RTB_MMA_CP * TWP;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int n_wtiles = B / TW;
    const long long n_tiles_total = (long long)S * n_wtiles;
    const long long my_tiles = (n_tiles_total - blockIdx.x + gridDim.x - 1) / gridDim.x;
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // channel rows beyond C stay zero in every stage (the bulk copies never touch them)
    for (int st = 0; st < stages; ++st)
        for (int i = C * TWP + threadIdx.x; i < tile_elems; i += 128) tiles[(size_t)st * tile_elems + i] = 0.f;
    __syncthreads();

    const long long n_seq = (long long)n_jchunks * my_tiles;  // copies: the tiles once per row chunk
    auto issue = [&](long long seq) {  // warp 0: one bulk copy per channel row of copy number `seq`
        const long long tile = blockIdx.x + (seq % my_tiles) * gridDim.x;
        const int st = (int)(seq % stages);
        const long long s = tile / n_wtiles;
        const int t0 = (int)(tile % n_wtiles) * TW;
        if (lane == 0) mbar_expect_tx(&full_bar[st], (unsigned)(C * TW * sizeof(float)));
        __syncwarp();
        const float *src = x + ((size_t)s * C) * B + t0;
        float *dst = tiles + (size_t)st * tile_elems;
        for (int c = lane; c < C; c += 32)
            bulk_copy_g2s(dst + (size_t)c * TWP, src + (size_t)c * B, (unsigned)(TW * sizeof(float)), &full_bar[st]);
    };
    if (warp == 0)
        for (int i = 0; i < stages && i < n_seq; ++i) issue(i);

    const int n_tsub = TW / 64;  // warp tiles are 32 samples wide: 2 * n_tsub of them per tile
    for (int jc = 0; jc < n_jchunks; ++jc) {
        // B fragments of this row chunk: b0 = R[n0+g][k0+t4], b1 = R[n0+g][k0+t4+4], hi and lo
        uint32_t bh[4][4][2], bl[4][4][2];
        {
            const float *rh = rsp + (size_t)jc * 2 * RTB_MMA_JP * RTB_MMA_CP, *rl = rh + RTB_MMA_JP * RTB_MMA_CP;
#pragma unroll
            for (int kt = 0; kt < 4; ++kt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int o = (nt * 8 + g) * RTB_MMA_CP + kt * 8 + t4 + 4 * h;
                        bh[kt][nt][h] = __float_as_uint(__ldg(rh + o));
                        bl[kt][nt][h] = __float_as_uint(__ldg(rl + o));
                    }
        }
        // with more than one row chunk the tiles are streamed once per chunk (jc outer): the pipeline
        // below simply runs n_jchunks times over this CTA's tiles
        for (long long it0 = 0; it0 < my_tiles; ++it0) {
            const long long it = (long long)jc * my_tiles + it0;  // position in the copy sequence
            const int st = (int)(it % stages);
            mbar_wait(&full_bar[st], (unsigned)((it / stages) & 1));
            const long long tile = blockIdx.x + it0 * gridDim.x;
            const long long s = tile / n_wtiles;
            const int tbase = (int)(tile % n_wtiles) * TW;
            const float *xs = tiles + (size_t)st * tile_elems;
            for (int th = warp; th < 2 * n_tsub; th += 4) {  // warp tile: 32 samples x 32 rows
                float acc[2][4][4];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                        for (int r = 0; r < 4; ++r) acc[mt][nt][r] = 0.f;
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    const int m0 = th * 32 + mt * 16;
#pragma unroll
                    for (int kt = 0; kt < 4; ++kt) {
                        // A fragment (16 samples x 8 channels): a0 (g, t4), a1 (g+8, t4), a2 (g, t4+4), a3 (g+8, t4+4)
                        const float *ap = xs + (size_t)(kt * 8 + t4) * TWP + m0 + g;
                        uint32_t ah[4], al[4];
                        split_tf32_rna(ap[0], ah[0], al[0]);
                        split_tf32_rna(ap[8], ah[1], al[1]);
                        split_tf32_rna(ap[4 * TWP], ah[2], al[2]);
                        split_tf32_rna(ap[4 * TWP + 8], ah[3], al[3]);
#pragma unroll
                        for (int nt = 0; nt < 4; ++nt) {
                            mma_tf32_16x8x8(acc[mt][nt], al, bh[kt][nt][0], bh[kt][nt][1]);
                            mma_tf32_16x8x8(acc[mt][nt], ah, bl[kt][nt][0], bl[kt][nt][1]);
                            mma_tf32_16x8x8(acc[mt][nt], ah, bh[kt][nt][0], bh[kt][nt][1]);
                        }
                    }
                }
                // D fragment: c0 (sample g, row 2 t4), c1 (g, 2 t4 + 1), c2 (g + 8, 2 t4), c3 (g + 8, 2 t4 + 1)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const int j0 = jc * RTB_MMA_JP + nt * 8 + 2 * t4;
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        float *zp = zcur + ((size_t)s * J + j0) * B + tbase + th * 32 + mt * 16 + g;
                        if (j0 < J) { zp[0] = acc[mt][nt][0]; zp[8] = acc[mt][nt][2]; }
                        if (j0 + 1 < J) { zp[B] = acc[mt][nt][1]; zp[B + 8] = acc[mt][nt][3]; }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[st]);
            if (warp == 0 && it + stages < n_seq) {
                mbar_wait(&empty_bar[st], (unsigned)((it / stages) & 1));
                issue(it + stages);
            }
        }
    }
}

}  // namespace rtb
