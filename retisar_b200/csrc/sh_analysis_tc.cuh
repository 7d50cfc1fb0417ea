// K-A on the 5th-generation tensor cores (tcgen05, accumulator in TMEM) for wide arrays / high SH
// orders, where the analysis GEMM  z[j][t] = sum_c Rm[j][c] x[c][t]  is FP32-pipe bound on SIMT
// (order 8, 110 channels: 23 flop/B).  Included by sh_kernels.cuh.
//
//   D[t][j] (TMEM, fp32)  =  A[t][c] (smem, K-major)  x  B[j][c]^T (smem, K-major)
//   M = 128 samples of one stream, N = rows padded to 16, K = channels in chunks of KC, kind::tf32.
//
// Plain TF32 (10-bit mantissa) would break the 1e-5 parity bar, so both operands are split
// a = a_hi + a_lo (hi = a rounded to TF32, lo = a - hi, exact in fp32) and three MMAs accumulate
// hi*hi + lo*hi + hi*lo in the fp32 accumulator; measured max error 2e-6 relative
// (tools/tc_probe.cu).  The analysis matrix is split and laid out on the host; the input tile is
// split on the fly while it is scattered from coalesced global loads into the un-swizzled
// core-matrix layout [c/4][t/8][t%8][c%4] (a thread's four channels form one 16-byte run, a warp
// writes 512 contiguous bytes).  MN-major operands would avoid that transposition but produced
// zeros with un-swizzled TF32 descriptors on this part (tools/tc_probe.cu, variant 0), so both
// operands are K-major.
//
// Warp-specialised, persistent (one CTA per SM, 13 warps):
//   warps 0-3 / 4-7  two loader groups alternating over the K-chunks, each filling its own
//                    operand stage: row reads (the group's next chunk is requested before the
//                    current one is processed), TF32 split, scatter, fence.proxy.async
//   warp 12          one thread issues the MMAs of a chunk and commits the stage back
//   warps 8-11       epilogue: TMEM lane quarter -> registers -> analysis rows (coalesced 128 B)
// A ring of 2-4 operand stages and two TMEM accumulators let the three roles overlap (with two
// stages the loaders could only start on a stage once its MMAs had retired: a clock64 trace showed
// the MMA issuer waiting ~3000 of the 7000 cycles per tile for them).
//
// Tile rows are the flattened (stream, sample) index, so blocks shorter than 128 samples put 2 or 4
// streams into one M = 128 tile (low-latency 64-sample blocks, config c4).  When the split analysis
// matrix does not fit in shared memory next to the operand stages (434 channels x 169 rows: 630 KB)
// its K-chunks are streamed with cp.async.bulk into the same two stages as the input (STREAM_B):
// the chunk of a stage is requested by the loader group that fills the stage, completion is counted
// on the stage's `full` mbarrier together with the loader warps' arrivals.
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

__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

// -DRTB_TRACE_TC: CTA 0 logs clock64 stamps per role and tile (tools/tc_trace.py)
#ifdef RTB_TRACE_TC
__device__ long long g_tc_trace[4096];
#define TC_STAMP(role, it, k) do { if (blockIdx.x == 0 && (it) < 100) g_tc_trace[(role) * 1000 + (it) * 8 + (k)] = clock64(); } while (0)
#else
#define TC_STAMP(role, it, k) do {} while (0)
#endif

#define RTB_TC_M 128
#define RTB_TC_MAX_KC 64   // channels per pipeline chunk (register staging of the loaders)
#define RTB_TC_THREADS 416
#define RTB_TC_MAX_ST 4     // operand stages (ring): the loaders run up to nst - 1 chunks ahead of the MMAs

template <int KC, bool STREAM_B>
__global__ void __launch_bounds__(RTB_TC_THREADS, 1)
sh_analysis_tc_kernel(const float *__restrict__ x,     // [S][C][B]
                      const float *__restrict__ bsrc,  // [2][Kpad/4][Npad/8][8][4]: hi then lo
                      float *__restrict__ zcur,        // [S][J][B]
                      int S, int C, int J, int B, int Kpad, int Npad, int acc_cols, int nst) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *a_stage = reinterpret_cast<float *>(smem_raw);   // [nst stages][hi, lo][KC/4][16][8][4]
    // resident: [Kpad/4][Npad/8][8][4] hi, then lo; streamed: [2 stages][hi, lo][KC/4][Npad/8][8][4]
    float *b_hi = a_stage + (size_t)nst * 2 * KC * RTB_TC_M;
    float *b_lo = b_hi + (size_t)Kpad * Npad;
    const size_t bchunk_floats = (size_t)KC * Npad;         // one K-chunk of hi (or lo)
    __shared__ __align__(8) uint64_t full_bar[RTB_TC_MAX_ST], empty_bar[RTB_TC_MAX_ST], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (!STREAM_B)
        for (int i = tid; i < Kpad * Npad / 2; i += RTB_TC_THREADS)  // hi and lo halves are contiguous
            reinterpret_cast<float4 *>(b_hi)[i] = __ldg(reinterpret_cast<const float4 *>(bsrc) + i);
    if (tid == 0) {
        for (int i = 0; i < RTB_TC_MAX_ST; ++i) {
            // one arrival per loader warp of the stage's group (+ the expect_tx arrival of the B chunk)
            mbar_init(&full_bar[i], STREAM_B ? 5 : 4);
            mbar_init(&empty_bar[i], 1);   // tcgen05.commit
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);   // tcgen05.commit
            mbar_init(&tempty_bar[i], 4);  // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(2 * acc_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // B operand written by generic stores
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    // tile = 128 consecutive rows of the flattened [S][B] sample index
    const long long n_rows = (long long)S * B;
    const long long n_tiles = (n_rows + RTB_TC_M - 1) / RTB_TC_M;
    const long long my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const int n_chunks = Kpad / KC;
    const long long my_chunks = my_tiles * n_chunks;
    const size_t stage_floats = (size_t)2 * KC * RTB_TC_M;

    if (warp < 8) {
        // ===== loaders: thread = sample t of the tile.  Group g = warp / 4 takes the chunks with
        // n % 2 == g, i.e. always fills stage g.
        const int t_in = tid & 127, grp_l = warp >> 2;
        auto issue_loads = [&](long long n, float (&xr)[KC]) {
            const long long tile = blockIdx.x + (n / n_chunks) * gridDim.x;
            const int h = (int)(n % n_chunks);
            const long long g = tile * RTB_TC_M + t_in;     // flattened (stream, sample) row
            const bool row_ok = g < n_rows;
            const long long s = row_ok ? g / B : 0;
            const int t = row_ok ? (int)(g % B) : 0;
            const float *xp = x + ((size_t)s * C + (size_t)h * KC) * B + t;
            const int cmax = row_ok ? C - h * KC : 0;
#pragma unroll
            for (int i = 0; i < KC; ++i) {  // predicated: channels beyond C (and rows beyond S*B) read as zero
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.lt.s32 p, %2, %3;\n\tmov.f32 %0, 0f00000000;\n\t"
                    "@p ld.global.nc.f32 %0, [%1];\n\t}"
                    : "=f"(xr[i]) : "l"(xp), "r"(i), "r"(cmax));
                xp += B;
            }
        };
        auto process = [&](long long n, float (&xr)[KC]) {
            const int st = (int)(n % nst);
            if (t_in == 0) TC_STAMP(grp_l, (int)(n >> 1), 0);
            mbar_wait(&empty_bar[st], (unsigned)(((n / nst) & 1) ^ 1));
            if (t_in == 0) TC_STAMP(grp_l, (int)(n >> 1), 1);
            if (STREAM_B && t_in == 0) {  // this chunk's rows of the split analysis matrix -> stage st
                const int h = (int)(n % n_chunks);
                const unsigned bytes = (unsigned)(bchunk_floats * sizeof(float));
                float *bst = b_hi + (size_t)st * 2 * bchunk_floats;
                mbar_expect_tx(&full_bar[st], 2 * bytes);
                bulk_copy_g2s(bst, bsrc + (size_t)h * bchunk_floats, bytes, &full_bar[st]);
                bulk_copy_g2s(bst + bchunk_floats, bsrc + (size_t)Kpad * Npad + (size_t)h * bchunk_floats, bytes,
                              &full_bar[st]);
            }
            float *ahi = a_stage + st * stage_floats, *alo = ahi + (size_t)KC * RTB_TC_M;
#pragma unroll
            for (int u = 0; u < KC / 4; ++u) {
                float hi[4], lo[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) split_tf32(xr[4 * u + i], hi[i], lo[i]);
                const size_t off = (((size_t)u * (RTB_TC_M / 8) + (t_in >> 3)) * 8 + (t_in & 7)) * 4;
                *reinterpret_cast<float4 *>(ahi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4 *>(alo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> tensor core reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_bar[st]);
            if (t_in == 0) TC_STAMP(grp_l, (int)(n >> 1), 2);
        };
        float xa[KC], xb[KC];
        long long n = grp_l;  // this group's chunks: grp_l, grp_l + 2, ...
        if (n < my_chunks) issue_loads(n, xa);
        for (; n < my_chunks; n += 4) {
            if (n + 2 < my_chunks) issue_loads(n + 2, xb);
            process(n, xa);
            if (n + 2 < my_chunks) {
                if (n + 4 < my_chunks) issue_loads(n + 4, xa);
                process(n + 2, xb);
            }
        }
    } else if (warp == 12) {
        // ===== MMA issuer (one thread): hi*hi + lo*hi + hi*lo per chunk, accumulating in TMEM
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(Npad >> 3) << 17) |
                                   ((uint32_t)(RTB_TC_M >> 4) << 24);  // D f32, A/B tf32, K-major, N, M
            const uint32_t lbo_a = (RTB_TC_M / 8) * 128, lbo_b = (Npad / 8) * 128, sbo = 128;
            const uint64_t step_a = (2 * lbo_a) >> 4, step_b = (2 * lbo_b) >> 4;
            long long n = 0;
            for (long long it = 0; it < my_tiles; ++it) {
                const int acc = (int)(it & 1);
                TC_STAMP(2, (int)it, 0);
                mbar_wait(&tempty_bar[acc], (unsigned)(((it >> 1) & 1) ^ 1));
                TC_STAMP(2, (int)it, 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem + acc * acc_cols;
                for (int h = 0; h < n_chunks; ++h, ++n) {
                    const int st = (int)(n % nst);
                    mbar_wait(&full_bar[st], (unsigned)((n / nst) & 1));
                    TC_STAMP(2, (int)it, 2 + h);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    // descriptors differ only in the 14-bit start-address field: build them once per
                    // chunk and step the field by a constant (no carry: shared memory is < 256 KB)
                    const uint32_t ahi = smem_u32(a_stage + st * stage_floats);
                    const uint32_t alo = ahi + (uint32_t)(KC * RTB_TC_M * sizeof(float));
                    const uint64_t da_hi = umma_desc(ahi, lbo_a, sbo), da_lo = umma_desc(alo, lbo_a, sbo);
                    // this chunk's k-range of B: inside the resident matrix, or the stage's own copy
                    const uint32_t bhi_a = STREAM_B ? smem_u32(b_hi + (size_t)st * 2 * bchunk_floats)
                                                    : smem_u32(b_hi) + (uint32_t)h * (KC / 4) * lbo_b;
                    const uint32_t blo_a = STREAM_B ? bhi_a + (uint32_t)(bchunk_floats * sizeof(float))
                                                    : smem_u32(b_lo) + (uint32_t)h * (KC / 4) * lbo_b;
                    const uint64_t db_hi = umma_desc(bhi_a, lbo_b, sbo);
                    const uint64_t db_lo = umma_desc(blo_a, lbo_b, sbo);
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
                        const uint64_t da0 = pass == 1 ? da_lo : da_hi;
                        const uint64_t db0 = pass == 2 ? db_lo : db_hi;
#pragma unroll
                        for (int ks = 0; ks < KC / 8; ++ks)
                            umma_tf32(d_tmem, da0 + ks * step_a, db0 + ks * step_b, idesc,
                                      (h | pass | ks) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[st]);  // the stage is free once these MMAs have read it
                }
                umma_commit(&tfull_bar[acc]);
                TC_STAMP(2, (int)it, 6);
            }
        }
    } else {
        // ===== epilogue: warp w drains TMEM lanes 32*(w%4).. (rows t) to the analysis rows z[j][t]
        const int q = warp & 3;
        for (long long it = 0; it < my_tiles; ++it) {
            const long long tile = blockIdx.x + it * gridDim.x;
            const long long g = tile * RTB_TC_M + q * 32 + lane;  // this lane's (stream, sample) row
            const bool row_ok = g < n_rows;
            const long long s = row_ok ? g / B : 0;
            const int t = row_ok ? (int)(g % B) : 0;
            const int acc = (int)(it & 1);
            if (q == 0 && lane == 0) TC_STAMP(3, (int)it, 0);
            mbar_wait(&tfull_bar[acc], (unsigned)((it >> 1) & 1));
            if (q == 0 && lane == 0) TC_STAMP(3, (int)it, 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float *zp = zcur + ((size_t)s * J) * B + t;
            for (int cb = 0; cb < Npad; cb += 32) {
                uint32_t v[32];
                tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + acc * acc_cols + cb, v);
                const int nrow = row_ok ? J - cb : 0;  // rows of this chunk that exist
#pragma unroll
                for (int j = 0; j < 32; ++j) {  // predicated stores, pointer stepped by one row
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.s32 p, %2, %3;\n\t@p st.global.b32 [%0], %1;\n\t}"
                                 ::"l"(zp), "r"(v[j]), "r"(j), "r"(nrow) : "memory");
                    zp += B;
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            if (q == 0 && lane == 0) TC_STAMP(3, (int)it, 2);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(2 * acc_cols));
}

}  // namespace rtb
