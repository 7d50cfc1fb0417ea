// K-A on the 5th-generation tensor cores (tcgen05, accumulator in TMEM): the analysis GEMM
// z[j][t] = sum_c Rm[j][c] x[c][t]  for arrays of 32 channels and more.  Included by sh_kernels.cuh.
//
//   D[t][j] (TMEM, fp32)  =  A[t][c] (smem, K-major)  x  B[j][c]^T (smem, K-major)
//   M = 128 samples, N = rows padded to 16, K = channels in chunks of 32 (one operand stage), kind::tf32.
//
// Plain TF32 (10-bit mantissa) would break the 1e-5 parity bar, so both operands are split
// a = a_hi + a_lo (hi = a rounded to TF32, lo = a - hi, exact in fp32) and three MMAs accumulate
// hi*hi + lo*hi + hi*lo in the fp32 accumulator; measured max error 2e-6 relative
// (tools/tc_probe.cu).  The analysis matrix is split and laid out on the host; the input tile is
// split on the fly while it is transposed from 16-byte global loads into the un-swizzled
// core-matrix layout [c/4][t/8][t%8][c%4].  MN-major operands would avoid that transposition but
// produced zeros with TF32 descriptors on this part (tools/tc_probe_mn.cu), so both operands are
// K-major; the 128-byte-swizzled K-major layout issues at the same rate as the un-swizzled one
// (tools/tc_rate.cu: a K = 8 TF32 MMA takes max(48, N/2) cycles, i.e. for N <= 96 it is bound by
// the 4 KB A-operand fetch from shared memory).
//
// Warp-specialised, persistent (one CTA per SM, 2 * RTB_TC_LG + 5 = 13 warps):
//   warps 0-7   loaders: RTB_TC_LG = four pairs of warps, each with its own full / empty mbarriers, running out
//               of phase -- 16-byte row loads, TF32 split, conflict-free transposed stores,
//               fence.proxy.async, arrive on the stage's `full` mbarrier
//   warp 12     MMA issuer: warp-uniform loop, one elected lane issues the MMAs of a stage and
//               commits the stage back (`empty`) / the accumulator to the epilogue (`tfull`)
//   warps 8-11  epilogue: TMEM lane quarter -> registers -> analysis rows (coalesced 128 B)
// Two TMEM accumulators let the epilogue of a tile overlap the MMAs of the next.
//
// What the clock64 traces (tools/tc_trace.py) and the SASS showed on the way here (c5, 8 192 streams:
// 503 -> 353 us; c2, 10 240 streams: 239 -> 106 us, the mma.sync kernel takes 132 us):
//   * every LDG of a loader loop sits on ONE scoreboard, so a register prefetch ring never overlapped
//     anything: the first use of the oldest buffer waited for the youngest load (one memory round trip
//     per stage, whatever the ring depth, load width, cache hint or tile shape).  The loaders now
//     run a plain load / wait / split / store cycle and overlap ACROSS warp groups instead.
//   * fence.proxy.async is lowered to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC; the MEMBAR waits for the
//     thread's loads in flight (second reason the prefetch ring was useless).
//   * with `if (lane == 0)` around the issue loop every tcgen05.mma was wrapped in an elect / R2UR
//     broadcast loop (~110 cycles per MMA against the MMA's own 49): the issue loop must be
//     warp-uniform with the warp index taken from a shuffle.
//
// Tile rows are the flattened (stream, sample) index, so blocks shorter than 128 samples put 2 or 4
// streams into one M = 128 tile (low-latency 64-sample blocks, config c4).  When the split analysis
// matrix does not fit in shared memory next to the operand stages (434 channels x 169 rows: 630 KB)
// its K-chunks are streamed with cp.async.bulk into the stages next to the input (STREAM_B): the
// chunk of a stage is requested by the loader group that fills the stage, completion is counted on
// the stage's `full` mbarrier together with the loader warps' arrivals.
#pragma once
#include "rtb_common.cuh"
#include "rtb_async.cuh"

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
__device__ __forceinline__ uint32_t elect_one_sync() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred;
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

// -DRTB_TRACE_TC: CTA 0 logs clock64 stamps per role (tools/tc_trace.py): role 0 = loader warp 0
// (indexed by chunk), role 2 = MMA issuer, role 3 = epilogue warp 8 (indexed by tile)
#ifdef RTB_TRACE_TC
static __device__ long long g_tc_trace[4096];
#define TC_STAMP(role, it, k) do { if (blockIdx.x == 0 && (it) < 100) g_tc_trace[(role) * 1000 + (it) * 8 + (k)] = clock64(); } while (0)
#else
#define TC_STAMP(role, it, k) do {} while (0)
#endif

#define RTB_TC_M 128
#define RTB_TC_KC 32       // channels per pipeline chunk (operand stage): 16 per warp of a loader pair
#ifndef RTB_TC_LG
// loader groups (pairs of warps), each with 16 KB of loads in flight.  Measured with 6 groups (17 warps,
// 96 KB in flight): c5 343 vs 356 us, c2 10 240 streams 99 vs 106 us, but c2 1 024 streams 22.3 vs
// 21.8 us and c4 61.3 vs 59.9 us; 8 groups (80 registers, spills): slower everywhere.  The launch
// picks RTB_TC_LG_BIG groups when every CTA has many tiles (sh_plan.cu: launch_analysis).
#define RTB_TC_LG 4
#endif
#define RTB_TC_LG_BIG 6     // for launches with many tiles per CTA (chosen at launch)
#define RTB_TC_THREADS_OF(LG) (32 * (2 * (LG) + 5))   // loaders, 4 epilogue warps, MMA issuer
#define RTB_TC_THREADS RTB_TC_THREADS_OF(RTB_TC_LG)
#define RTB_TC_MAX_ST 4    // operand stages (ring)
// Where the generic-proxy -> async-proxy fence between the loaders' st.shared and the tensor core's
// operand reads sits.  nvcc lowers fence.proxy.async to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC, and the
// MEMBAR waits for everything the executing thread has in flight:
//   1 (default) in the loader warps after their stores: their loads have all been consumed by then
//     (no register prefetch, see the loader comment), so only the stores drain;
//   0 in the MMA-issuing thread after it has acquired the stage through the `full` mbarrier (also on
//     the causality path between the stores and the reads): measured slower, the MEMBAR drains the
//     MMAs of the previous stage before the next ones are issued (112 instead of 49 cycles per MMA).
#ifndef RTB_TC_FENCE_WRITER
#define RTB_TC_FENCE_WRITER 1
#endif

// 16 bytes of one channel row (4 consecutive samples), zero when the predicate is off (the
// predicate is folded into the address).
static __device__ float4 g_tc_zero4;  // zero-initialised
__device__ __forceinline__ float4 ldg_nc_v4_pred(const float *p, bool pred) {
    return __ldg(reinterpret_cast<const float4 *>(pred ? p : reinterpret_cast<const float *>(&g_tc_zero4)));
}

// RING: the input rows form a ring (the mono convolver's delay line): channel c < kmod is read from
// row (c + krot) % kmod.  A template flag, not a run-time test: the extra address arithmetic in the
// loaders cost the analysis launches 2 us of 20 (c2, 1 024 streams) when it was one.
template <bool STREAM_B, int LG = RTB_TC_LG, bool RING = false>
__global__ void __launch_bounds__(RTB_TC_THREADS_OF(LG), 1)
sh_analysis_tc_kernel(const float *__restrict__ x,     // [S][C][B]
                      const float *__restrict__ bsrc,  // [2][Kpad/4][Npad/8][8][4]: hi then lo
                      float *__restrict__ zcur,        // [S][J][B]
                      int S, int C, int J, int B, int Kpad, int Npad, int acc_cols, int nst,
                      long long bmat_stride,   // STREAM_B: floats between the matrices of consecutive streams (0: one matrix)
                      int krot = 0, int kmod = 0) {  // ring of input rows: channel c < kmod is read from row (c + krot) % kmod
    constexpr int KC = RTB_TC_KC;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *a_stage = reinterpret_cast<float *>(smem_raw);   // [nst stages][hi, lo][KC/4][16][8][4]
    // resident: [Kres/4][Npad/8][8][4] hi, then lo (Kres = C rounded up to 8: the k-steps that are
    // issued); streamed: [nst stages][hi, lo][KC/4][Npad/8][8][4]
    const int Kres = (C + 7) / 8 * 8;
    float *b_hi = a_stage + (size_t)nst * 2 * KC * RTB_TC_M;
    float *b_lo = b_hi + (size_t)Kres * Npad;
    const size_t bchunk_floats = (size_t)KC * Npad;         // one K-chunk of hi (or lo)
    __shared__ __align__(8) uint64_t full_bar[LG], empty_bar[LG], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_s;
    // warp index through a shuffle: provably warp-uniform, so that the role branches are uniform and
    // the MMA issuer's descriptors live in uniform registers (with `tid >> 5` and an `if (lane == 0)`
    // around the issue loop ptxas wrapped every tcgen05.mma into an elect / R2UR broadcast loop:
    // ~22 instructions and ~110 cycles per MMA, twice the MMA's own 49 cycles)
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

    if (!STREAM_B)
        for (int i = tid; i < Kres * Npad / 4; i += RTB_TC_THREADS_OF(LG)) {
            reinterpret_cast<float4 *>(b_hi)[i] = __ldg(reinterpret_cast<const float4 *>(bsrc) + i);
            reinterpret_cast<float4 *>(b_lo)[i] = __ldg(reinterpret_cast<const float4 *>(bsrc + (size_t)Kpad * Npad) + i);
        }
    if (tid == 0) {
        for (int i = 0; i < LG; ++i) {
            // one arrival per warp of the loader group (+ the expect_tx arrival of the B chunk)
            mbar_init(&full_bar[i], 2 + (STREAM_B ? 1 : 0));
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

    // tile = 128 consecutive rows of the flattened [S][B] sample index.  (256-row tiles that take whole
    // 1 KB channel rows at once, as two M = 128 halves with an accumulator each, were measured equal
    // or slightly slower: c5 358.9 vs 352.7 us, c2 106.0 vs 107.4 us.)
    const long long n_rows = (long long)S * B;
    const long long n_tiles = (n_rows + RTB_TC_M - 1) / RTB_TC_M;
    const long long my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const int n_chunks = Kpad / KC;
    const long long my_units = my_tiles * n_chunks;   // unit = (tile, chunk): one operand stage
    const size_t stage_floats = (size_t)2 * KC * RTB_TC_M;
    const int log2B = 31 - __clz(B);  // block lengths are powers of two

    if (warp < 2 * LG) {
        // ===== loaders.  A stage (32 channels x 128 rows) is filled by a group of warps (a PAIR with 16
        // channels each when there are four stages); group p takes the units n = p, p + nst, ...  Lane l owns rows 4l .. 4l+3 of the stage (one
        // 16-byte load per channel: a warp request covers the 512 contiguous bytes of a channel row).
        // ptxas puts every LDG of this loop on one scoreboard, so a warp cannot overlap its own
        // loads with its own stores (the first use of any loaded register waits for all loads in
        // flight; measured: a register prefetch ring ran at one memory round trip per stage whatever
        // its depth).  Each warp therefore runs a plain issue-16-loads / wait / split / store cycle
        // and the four pairs run out of phase: while one pair splits and stores, the other three
        // have their 16 KB in flight.
        // The 4 x 4 register block of a channel quad is transposed on the way to shared memory:
        // step s stores row 4l + (s + rot) % 4 with rot = (l / 2) % 4, so that the eight lanes of a
        // 128-bit store phase hit eight different 16-byte bank groups (rows 4l + s for all lanes
        // would be a 4-way conflict).
        // Four loader groups (pairs of warps, 16 channels per warp) whatever the number of physical
        // stages: unit n (in MMA order) belongs to group n % 4 and lands in stage n % nst.  Every
        // group has its own full / empty mbarrier pair (a barrier shared by two groups would be
        // waited on every other phase and its parity misread): after unit n the MMA issuer commits
        // to the `empty` barrier of group (n + nst) % 4, the next user of that stage.  With two
        // stages (streamed B operand) four groups still keep 64 KB of loads in flight.
        constexpr int qpw = 4;                             // channel quads per warp
        const int pair = warp >> 1, sub = warp & 1;        // loader group, warp in the group
        const int rot = (lane >> 1) & 3;
        const bool r1 = rot & 1, r2 = rot & 2;
        for (long long n = pair; n < my_units; n += LG) {
            const long long it = n / n_chunks;
            const int h = (int)(n - it * n_chunks);
            const long long g = ((long long)blockIdx.x + it * gridDim.x) * RTB_TC_M + 4 * lane;
            const bool row_ok = g < n_rows;
            const long long s = row_ok ? (g >> log2B) : 0;
            const int t = row_ok ? (int)(g & (B - 1)) : 0;
            const int c0 = h * KC + 4 * qpw * sub;
            const float *xp = x + ((size_t)s * C + c0) * B + t;
            // channels beyond C or beyond this warp's share (and rows beyond S*B) read as zero
            const int cmax = row_ok ? min(C - c0, 4 * qpw) : 0;
            float4 v[4][4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if constexpr (RING) {
                        int off = 4 * q + i;  // channel c0 + off: its row rotates
                        if (c0 + off < kmod) {
                            int r = c0 + off + krot;
                            if (r >= kmod) r -= kmod;
                            off = r - c0;
                        }
                        v[q][i] = ldg_nc_v4_pred(xp + (long long)off * B, 4 * q + i < cmax);
                    } else {
                        v[q][i] = ldg_nc_v4_pred(xp + (size_t)(4 * q + i) * B, 4 * q + i < cmax);
                    }
                }
            const int st = (int)(n % nst);
            // k-th unit of this group: its stage was last used by unit n - nst, whose completion is
            // commit number k (groups that start on a fresh stage) or k + 1 (the others) on this
            // group's `empty` barrier
            const unsigned k = (unsigned)(n / LG);
            const unsigned par = (pair < nst) ? ((k & 1) ^ 1) : (k & 1);
            if (lane == 0 && sub == 0) TC_STAMP(0, (int)n, 0);
            mbar_wait(&empty_bar[pair], par);
            if (lane == 0 && sub == 0) TC_STAMP(0, (int)n, 1);
            if (STREAM_B && sub == 0 && lane == 0) {  // this chunk's rows of the split analysis matrix -> stage st
                const unsigned bytes = (unsigned)(bchunk_floats * sizeof(float));
                float *bst = b_hi + (size_t)st * 2 * bchunk_floats;
                const float *bs = bsrc + (size_t)s * bmat_stride;  // lane 0: first row of the tile
                mbar_expect_tx(&full_bar[pair], 2 * bytes);
                bulk_copy_g2s(bst, bs + (size_t)h * bchunk_floats, bytes, &full_bar[pair]);
                bulk_copy_g2s(bst + bchunk_floats, bs + (size_t)Kpad * Npad + (size_t)h * bchunk_floats, bytes,
                              &full_bar[pair]);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (q < qpw && c0 + 4 * q < Kres) {  // quads beyond the issued k-steps are never read
                    float *ahi = a_stage + st * stage_floats + (qpw * sub + q) * (4 * RTB_TC_M);
                    float *alo = ahi + (size_t)KC * RTB_TC_M;
                    float w[4][4];  // w[i][sidx] = component (sidx + rot) % 4 of channel i
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float a1 = r1 ? v[q][i].y : v[q][i].x, b1 = r1 ? v[q][i].z : v[q][i].y;
                        const float c1 = r1 ? v[q][i].w : v[q][i].z, d1 = r1 ? v[q][i].x : v[q][i].w;
                        w[i][0] = r2 ? c1 : a1; w[i][1] = r2 ? d1 : b1;
                        w[i][2] = r2 ? a1 : c1; w[i][3] = r2 ? b1 : d1;
                    }
#pragma unroll
                    for (int sidx = 0; sidx < 4; ++sidx) {
                        float hi[4], lo[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) split_tf32(w[i][sidx], hi[i], lo[i]);
                        const int off = (4 * lane + ((sidx + rot) & 3)) * 4;
                        *reinterpret_cast<float4 *>(ahi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<float4 *>(alo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                    }
                }
            }
#if RTB_TC_FENCE_WRITER
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> tensor core reads
#endif
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_bar[pair]);
            if (lane == 0 && sub == 0) TC_STAMP(0, (int)n, 2);
        }
    } else if (warp == 2 * LG + 4) {
        // ===== MMA issuer: the whole warp runs the (uniform) loop, one elected lane issues
        // hi*hi + lo*hi + hi*lo per stage, accumulating in TMEM
        {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(Npad >> 3) << 17) |
                                   ((uint32_t)(RTB_TC_M >> 4) << 24);  // D f32, A/B tf32, K-major, N, M
            const uint32_t lbo_a = (RTB_TC_M / 8) * 128, lbo_b = (Npad / 8) * 128, sbo = 128;
            const uint64_t step_a = (2 * lbo_a) >> 4, step_b = (2 * lbo_b) >> 4;
            int st = 0;        // physical stage of the unit
            long long mm = 0;  // unit counter: loader group mm % 4, its phase (mm / 4) % 2
            for (long long it = 0; it < my_tiles; ++it) {
                const int acc = (int)(it & 1);
                if (lane == 0) TC_STAMP(2, (int)it, 0);
                mbar_wait(&tempty_bar[acc], (unsigned)(((it >> 1) & 1) ^ 1));
                if (lane == 0) TC_STAMP(2, (int)it, 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int h = 0; h < n_chunks; ++h) {
                    const int ksteps = min(KC / 8, (C - h * KC + 7) / 8);  // the last chunk may be short
                    {
                        const uint32_t d_tmem = tmem + acc * acc_cols;
                        mbar_wait(&full_bar[mm % LG], (unsigned)((mm / LG) & 1));
                        if (h == 0 && lane == 0) TC_STAMP(2, (int)it, 2);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#if !RTB_TC_FENCE_WRITER
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // loaders' stores -> tensor core reads
#endif
                        // descriptors differ only in the 14-bit start-address field: build them once per
                        // stage and step the field by a constant (no carry: shared memory is < 256 KB)
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
                        if (elect_one_sync()) {
#pragma unroll
                            for (int pass = 0; pass < 3; ++pass) {
                                const uint64_t da0 = pass == 1 ? da_lo : da_hi;
                                const uint64_t db0 = pass == 2 ? db_lo : db_hi;
#pragma unroll
                                for (int ks = 0; ks < KC / 8; ++ks)
                                    if (ks < ksteps)
                                        umma_tf32(d_tmem, da0 + ks * step_a, db0 + ks * step_b, idesc,
                                                  (h | pass | ks) ? 1u : 0u);
                            }
                            // the stage is free once these MMAs have read it: tell its next user
                            umma_commit(&empty_bar[(mm + nst) % LG]);
                            if (h == n_chunks - 1) umma_commit(&tfull_bar[acc]);
                        }
                        __syncwarp();
                        ++mm;
                        if (++st == nst) st = 0;
                    }
                }
                if (lane == 0) TC_STAMP(2, (int)it, 6);
            }
        }
    } else {
        // ===== epilogue: warp w drains TMEM lanes 32*(w%4).. (rows t) to the analysis rows z[j][t]
        const int q = warp & 3;
        for (long long it = 0; it < my_tiles; ++it) {
            const long long tile = blockIdx.x + it * gridDim.x;
            const int acc = (int)(it & 1);
            if (q == 0 && lane == 0) TC_STAMP(3, (int)it, 0);
            mbar_wait(&tfull_bar[acc], (unsigned)((it >> 1) & 1));
            if (q == 0 && lane == 0) TC_STAMP(3, (int)it, 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            {
                const long long g = tile * RTB_TC_M + q * 32 + lane;  // this lane's (stream, sample) row
                const bool row_ok = g < n_rows;
                const long long s = row_ok ? (g >> log2B) : 0;
                const int t = row_ok ? (int)(g & (B - 1)) : 0;
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
