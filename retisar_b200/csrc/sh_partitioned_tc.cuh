// Row a-X on the tensor cores: the partition MAC of the P-partition SH-domain HRIR as a per-bin GEMM
// on tcgen05, plus the stream-minor ("SoA") data layout it needs.  Included by sh_kernels.cuh after
// sh_partitioned.cuh (ShPartParams, ShStep) and sh_analysis_tc.cuh (descriptor / TMEM helpers).
//
// Per frequency bin f the contraction
//     queue_cl[q][ear][f][s] += sum_j H[q][j][ear][f] * ( phase_j(cl, s) * Zt_j(s, f) )
// (q = partition, j = (signal, term) step, cl = current / previous yaw, s = stream) is a complex
// GEMM whose matrix H[f] is shared by all streams: M = streams, K = steps, N = partitions x ears.
// Embedded in real arithmetic (row [.. a_re, a_im ..] times [[Hr, Hi], [-Hi, Hr]]) it runs as
//     D[128 streams][N = 4P] (TMEM, fp32)  +=  A[128][K = 2 steps] (smem)  x  B[N][K]^T (smem)
// with kind::tf32 and both operands split hi + lo (three MMAs: hi*hi + lo*hi + hi*lo), exactly like
// the analysis GEMM (sh_analysis_tc.cuh), because plain TF32 would break the 1e-5 parity bar.
//
// The A operand is produced on the fly: the spectra Z are read ONCE (stream-minor layout, one
// coalesced 256-byte request per warp and step), rotated by the phases of the current and of the
// previous yaw (two A tiles from one load), split and stored straight into the K-major core-matrix
// layout -- a k-quad is the (re, im) pairs of two steps of one stream, so a lane's 16-byte store
// needs no transposition.  The B operand (H of the bin, split on the host) is streamed by
// cp.async.bulk next to the A tiles of the stage.  The epilogue adds the accumulators to the two
// output-side queues, which are stored [bin][slot][ear][stream] so that a warp's read-modify-write
// is 256 contiguous bytes.
//
// Persistent, warp-specialised, one CTA per SM (25 warps):
//   warps 0-15   producers: four groups of four warps, group g owns operand stage g; a warp covers
//                32 streams x 8 steps of a unit = (tile, K-chunk of 16 reals)
//   warp 24      MMA issuer (warp-uniform loop, one elected lane)
//   warps 16-23  epilogue: TMEM lane quarter (warp % 4) x (current | previous yaw) accumulator, with the
//                queue rows of the next batch requested ahead (see the epilogue)
// What bounds it (ncu, profiles/r02_ncu_full_c3_partition_tc_*): HBM latency x bytes in flight.  v1
// (serial load - wait - store chain of four batches per tile in the epilogue): 513 us for c3 / 4 096
// streams = 3.5 TB/s; v2 (8 producer groups of 2 warps, 4 epilogue warps): 690 us, epilogue chain;
// v3 (5 producer groups of 2 warps, 8 epilogue warps, prefetching epilogue): 801 us, producers.
// A tile = (128 streams, one bin); CTAs take contiguous runs of tiles ordered (stream tile, bin), so
// a CTA meets at most two stream tiles per launch and keeps both phase tables in shared memory.
#pragma once

namespace rtb {

#define RTB_PT_KC 16        // reals per K-chunk (8 complex steps)
#define RTB_PT_LG 4         // producer groups = operand stages
#define RTB_PT_PW 4         // warps per producer group (32 streams each)
#define RTB_PT_EW 8         // epilogue warps
#define RTB_PT_THREADS (32 * (RTB_PT_LG * RTB_PT_PW + RTB_PT_EW + 1))
#define RTB_PT_MAX_STILES_PER_LAUNCH 128  // tiles of 128 streams per launch: a CTA's run spans <= 2

struct ShPartTcParams {
    const float2 *zspecT;        // [N2][NPT][Spad] spectra of the complex signals, stream-minor
    const float *bsrc;           // [F][n_chunks_max][hi, lo][KC/4][Npad/8][8][4] split H operand
    const ShStep *steps;         // [nsteps_max]
    const float *yaw_deg, *rad_last_in;   // [S] (offset to the launch's first stream)
    float2 *qcurT, *qlastT;      // [F][P][2][Spad] output-side queues, stream-minor
    long long Spad;              // allocated streams per row (multiple of 128)
    int s0;                      // first stream of this launch (multiple of 128)
    int S;                       // streams of the plan that exist (rows >= S are padding)
    int n_stiles;                // stream tiles of this launch
    int F, N2, NPT, P, Npad, acc_cols;
    int n_chunks;                // K-chunks that hold steps of this block (order lowering shortens K)
    int n_chunks_max;            // K-chunks of the operand table
    int nsteps_cur, nsteps_last; // steps (prefix) taking part in the current / previous yaw sum
    int order_max, do_last, head_cur, head_last;
    int overwrite_tail;          // partition P-1 (the head the previous block consumed, not cleared) is stored, not added to
    float tracker_dir, src_azim16;
};

// TF32 split with the conversion instruction (round to nearest, ties away): hi has 10 mantissa bits,
// lo = v - hi is exact in fp32; one instruction less per value than the integer rounding trick
__device__ __forceinline__ void split_tf32_cvt(float v, float &hi, float &lo) {
    unsigned u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    hi = __uint_as_float(u);
    lo = v - hi;
}

__device__ __forceinline__ float2 ldg_stream_v2(const float2 *p) {  // streamed once: keep it out of L1
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}

static __global__ void __launch_bounds__(RTB_PT_THREADS, 1)
sh_partition_tc_kernel(const ShPartTcParams prm) {
    constexpr int KC = RTB_PT_KC, LG = RTB_PT_LG, PW = RTB_PT_PW, EW = RTB_PT_EW, M = 128;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // stage: A [cl][hi, lo][KC/4][128][4] floats (32 KB), then B [hi, lo][KC/4][Npad/8][8][4]
    const int Npad = prm.Npad;
    const size_t a_floats = (size_t)2 * 2 * KC * M, b_floats = (size_t)2 * KC * Npad;
    const size_t stage_floats = a_floats + b_floats;
    float *stages = reinterpret_cast<float *>(smem_raw);
    float2 *s_ph = reinterpret_cast<float2 *>(stages + LG * stage_floats);  // [2 tables][2 cl][order+1][128]
    __shared__ __align__(8) uint64_t full_bar[LG], empty_bar[LG], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ ShStep s_steps[RTB_MAC_MAX_STEPS];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int n_ph = prm.order_max + 1;

    // this CTA's run of tiles; tile id = stile * F + bin
    const long long n_tiles = (long long)prm.n_stiles * prm.F;
    const long long tpc = (n_tiles + gridDim.x - 1) / gridDim.x;
    const long long t_begin = min(n_tiles, (long long)blockIdx.x * tpc);
    const long long t_end = min(n_tiles, t_begin + tpc);
    const long long my_tiles = t_end - t_begin;
    const int stile_first = (int)(t_begin / prm.F);
    const int nsteps = max(prm.nsteps_cur, prm.nsteps_last);

    for (int i = tid; i < nsteps; i += RTB_PT_THREADS) s_steps[i] = prm.steps[i];
    // phase tables (cos, sin)(m alpha) of the (at most two) stream tiles of the run
    for (int i = tid; i < 2 * M; i += RTB_PT_THREADS) {
        const int tab = i / M, r = i % M;
        const int s = prm.s0 + (stile_first + tab) * M + r;  // stream index in the plan
        const bool ok = s < prm.S && my_tiles > 0 && stile_first + tab < prm.n_stiles;
        const float rad_c = ok ? yaw_deg_to_rad16(prm.yaw_deg[s], prm.tracker_dir, prm.src_azim16) : 0.f;
        const float rad_l = ok ? prm.rad_last_in[s] : 0.f;
        for (int m = 0; m < n_ph; ++m) {
            float sn, cs;
            sincosf((float)m * rad_c, &sn, &cs);
            s_ph[((tab * 2 + 0) * n_ph + m) * M + r] = make_float2(cs, sn);
            sincosf((float)m * rad_l, &sn, &cs);
            s_ph[((tab * 2 + 1) * n_ph + m) * M + r] = make_float2(cs, sn);
        }
    }
    if (tid == 0) {
        for (int i = 0; i < LG; ++i) {
            mbar_init(&full_bar[i], PW + 1);  // the group's warps + the expect_tx arrival of the B chunk
            mbar_init(&empty_bar[i], 1);      // tcgen05.commit
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);      // tcgen05.commit
            mbar_init(&tempty_bar[i], EW);    // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(4 * prm.acc_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const int n_chunks = prm.n_chunks;
    const long long my_units = my_tiles * n_chunks;

    if (warp < LG * PW) {
        // ===== producers
        const int g = warp / PW, wq = warp % PW;
        const int r = 32 * wq + lane;               // row of the tile = stream inside the stream tile
        float *st_a = stages + g * stage_floats;
        float *st_b = st_a + a_floats;
        const size_t bunit = b_floats;              // floats of one (bin, chunk) operand block
        const int t_begin_i = (int)t_begin, n_units = (int)my_units;   // < 2^31: at most F tiles x 24 chunks per CTA
        for (int n = g, k = 0; n < n_units; n += LG, ++k) {
            const int tl = n / n_chunks, h = n - tl * n_chunks;
            const int tile = t_begin_i + tl;
            const int stile = tile / prm.F, f = tile - stile * prm.F;
            const int fm = (prm.N2 - f) & (prm.N2 - 1);
            const float2 *zrow = prm.zspecT + (size_t)stile * M + r;   // + (bin * NPT + p) * Spad
            const float2 *ph = s_ph + (size_t)(stile - stile_first) * 2 * n_ph * M + r;
            float2 z[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const int j = min(8 * h + jj, nsteps - 1);          // clamped: steps past the end get phase 0
                const ShStep stp = s_steps[j];
                z[jj] = ldg_stream_v2(zrow + ((size_t)(stp.term ? fm : f) * prm.NPT + stp.p) * prm.Spad);
            }
            mbar_wait_sleepy(&empty_bar[g], (unsigned)((k & 1) ^ 1));
            if (wq == 0 && lane == 0) {  // this unit's block of the split H operand -> stage
                const unsigned bytes = (unsigned)(bunit * sizeof(float));
                mbar_expect_tx(&full_bar[g], bytes);
                bulk_copy_g2s(st_b, prm.bsrc + ((size_t)f * prm.n_chunks_max + h) * bunit, bytes, &full_bar[g]);
            }
#pragma unroll
            for (int kq = 0; kq < KC / 4; ++kq) {
                float2 a[2][2];  // [cl][step of the quad]
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const int jj = 2 * kq + t, j = 8 * h + jj;
                    const ShStep stp = s_steps[min(j, nsteps - 1)];
                    float2 zz = z[jj];
                    if (stp.term) zz.y = -zz.y;
                    const int am = stp.m < 0 ? -stp.m : stp.m;
                    float2 pc = ph[(size_t)(0 * n_ph + am) * M], pl = ph[(size_t)(1 * n_ph + am) * M];
                    pc.y = stp.m < 0 ? pc.y : -pc.y;   // exp(-i m alpha) from (cos, sin)(|m| alpha)
                    pl.y = stp.m < 0 ? pl.y : -pl.y;
                    if (j >= prm.nsteps_cur) pc = make_float2(0.f, 0.f);
                    if (j >= prm.nsteps_last) pl = make_float2(0.f, 0.f);
                    a[0][t] = cmul(pc, zz);
                    a[1][t] = cmul(pl, zz);
                }
#pragma unroll
                for (int cl = 0; cl < 2; ++cl) {
                    if (cl == 1 && !prm.do_last) break;
                    float hi[4], lo[4];
                    split_tf32_cvt(a[cl][0].x, hi[0], lo[0]);
                    split_tf32_cvt(a[cl][0].y, hi[1], lo[1]);
                    split_tf32_cvt(a[cl][1].x, hi[2], lo[2]);
                    split_tf32_cvt(a[cl][1].y, hi[3], lo[3]);
                    float *dst = st_a + (size_t)cl * 2 * KC * M + (size_t)kq * (4 * M) + 4 * r;
                    *reinterpret_cast<float4 *>(dst) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4 *>(dst + KC * M) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> tensor core reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_bar[g]);
        }
    } else if (warp == LG * PW + EW) {
        // ===== MMA issuer
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(Npad >> 3) << 17) |
                               ((uint32_t)(M >> 4) << 24);  // D f32, A/B tf32, K-major, N, M
        const uint32_t lbo_a = (M / 8) * 128, lbo_b = (Npad / 8) * 128, sbo = 128;
        const uint64_t step_a = (2 * lbo_a) >> 4, step_b = (2 * lbo_b) >> 4;
        long long mm = 0;
        for (long long it = 0; it < my_tiles; ++it) {
            const int buf = (int)(it & 1);
            mbar_wait_sleepy(&tempty_bar[buf], (unsigned)(((it >> 1) & 1) ^ 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int h = 0; h < n_chunks; ++h, ++mm) {
                const int g = (int)(mm % LG);
                const int ksteps = min(KC / 8, (2 * nsteps - h * KC + 7) / 8);  // the last chunk may be short
                mbar_wait_sleepy(&full_bar[g], (unsigned)((mm / LG) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a0 = smem_u32(stages + g * stage_floats);
                const uint32_t b0 = a0 + (uint32_t)(a_floats * sizeof(float));
                const uint64_t db_hi = umma_desc(b0, lbo_b, sbo);
                const uint64_t db_lo = umma_desc(b0 + (uint32_t)(KC * Npad * sizeof(float)), lbo_b, sbo);
                if (elect_one_sync()) {
#pragma unroll
                    for (int cl = 0; cl < 2; ++cl) {
                        if (cl == 1 && !prm.do_last) break;
                        const uint32_t acl = a0 + (uint32_t)(cl * 2 * KC * M * sizeof(float));
                        const uint64_t da_hi = umma_desc(acl, lbo_a, sbo);
                        const uint64_t da_lo = umma_desc(acl + (uint32_t)(KC * M * sizeof(float)), lbo_a, sbo);
                        const uint32_t d_tmem = tmem + (uint32_t)((buf * 2 + cl) * prm.acc_cols);
#pragma unroll
                        for (int pass = 0; pass < 3; ++pass) {
                            const uint64_t da = pass == 1 ? da_lo : da_hi;
                            const uint64_t db = pass == 2 ? db_lo : db_hi;
#pragma unroll
                            for (int ks = 0; ks < KC / 8; ++ks)
                                if (ks < ksteps)
                                    umma_tf32(d_tmem, da + ks * step_a, db + ks * step_b, idesc,
                                              (h | pass | ks) ? 1u : 0u);
                        }
                    }
                    umma_commit(&empty_bar[g]);                      // the stage is free once these MMAs have read it
                    if (h == n_chunks - 1) umma_commit(&tfull_bar[buf]);
                }
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue: warp e drains TMEM lanes 32 * (e % 4) .. of the current (e < 4) or previous
        // yaw accumulator and adds them to the queue rows of its 32 streams.  The queue values of the
        // NEXT batch of rows (next column chunk / next tile) are requested before the warp waits for
        // the next accumulator -- their addresses do not depend on it -- so the read-modify-write
        // round trips overlap the MMAs instead of forming a serial chain per tile.
        constexpr int EB = 8;  // (partition, ear) rows per batch: 8 loads in flight per lane
        const int e = warp - LG * PW, quarter = warp & 3, cl = e / 4;
        float2 *qbase = cl ? prm.qlastT : prm.qcurT;
        const int head = cl ? prm.head_last : prm.head_cur;
        const bool idle = cl == 1 && !prm.do_last;
        // queue row of ((partition, ear) index) in the tile whose base pointer is `qt` (nullptr: padding row)
        auto tile_base = [&](long long it) -> float2 * {
            const int tile = (int)(t_begin + it);           // n_tiles < 2^31
            const int stile = tile / prm.F, f = tile - stile * prm.F;
            const int sl = stile * M + 32 * quarter + lane; // stream inside the launch
            if (it >= my_tiles || prm.s0 + sl >= prm.S) return nullptr;
            return qbase + (size_t)f * prm.P * 2 * prm.Spad + sl;
        };
        auto qptr = [&](float2 *qt, int idx) -> float2 * {
            const int part = idx >> 1;
            int slot = head + part;
            if (slot >= prm.P) slot -= prm.P;
            if (!qt || part >= prm.P) return nullptr;
            return qt + (size_t)(slot * 2 + (idx & 1)) * prm.Spad;
        };
        // The slot of partition P-1 holds the head the previous block consumed and nobody cleared: its
        // value is dropped where it is used.  It is still READ with the rest of its batch: skipping the
        // load (a select or a branch where the batch of 8 loads is issued) breaks the batch up and costs
        // 830 instead of 540 us per 4 096 streams; the gain is the tail kernel's clear (65 -> 30 us).
        auto fresh = [&](int idx) -> bool { return prm.overwrite_tail && (idx >> 1) == prm.P - 1; };
        float2 *qt_cur = idle ? nullptr : tile_base(0), *qt_next = nullptr;
        float2 old[EB];
        if (!idle && my_tiles > 0) {
#pragma unroll
            for (int c = 0; c < EB; ++c) {
                const float2 *q = qptr(qt_cur, c);
                if (q) old[c] = *q;
            }
        }
        for (long long it = 0; it < my_tiles; ++it) {
            const int buf = (int)(it & 1);
            if (!idle) qt_next = tile_base(it + 1);
            mbar_wait_sleepy(&tfull_bar[buf], (unsigned)((it >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int cb = 0; cb < Npad && !idle; cb += 32) {
                uint32_t v[32];
                tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((buf * 2 + cl) * prm.acc_cols + cb), v);
                const bool last_chunk = cb + 32 >= Npad;
                if (last_chunk) {  // the accumulator is in registers: release it to the MMA issuer
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty_bar[buf]);
                }
#pragma unroll
                for (int c0 = 0; c0 < 16; c0 += EB) {
                    // this batch was requested one batch ago: add, store (256 contiguous bytes per warp and row)
#pragma unroll
                    for (int c = 0; c < EB; ++c) {
                        float2 *q = qptr(qt_cur, cb / 2 + c0 + c);
                        if (q) {
                            const bool fr = fresh(cb / 2 + c0 + c);  // the value read is dropped, see `fresh`
                            *q = make_float2((fr ? 0.f : old[c].x) + __uint_as_float(v[2 * (c0 + c)]),
                                             (fr ? 0.f : old[c].y) + __uint_as_float(v[2 * (c0 + c) + 1]));
                        }
                    }
                    // request the next batch: rest of this chunk, the next chunk, or the next tile
                    const bool wrap = c0 + EB >= 16;
                    float2 *qt_n = (wrap && last_chunk) ? qt_next : qt_cur;
                    const int nidx = wrap ? (last_chunk ? 0 : (cb + 32) / 2) : cb / 2 + c0 + EB;
#pragma unroll
                    for (int c = 0; c < EB; ++c) {
                        const float2 *q = qptr(qt_n, nidx + c);
                        if (q) old[c] = *q;
                    }
                }
            }
            qt_cur = qt_next;
            if (idle) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[buf]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(4 * prm.acc_cols));
}

// Spectra of the complex SH signals in the stream-minor layout the kernel above reads:
// zspecT[bin][signal][stream].  One warp per stream, 8 streams per CTA; the spectra of a signal
// round are transposed through shared memory so that a bin row of the 8 streams leaves as 64
// contiguous bytes.
#ifndef RTB_SPEC_NW
#define RTB_SPEC_NW 8   // streams (warps) per CTA of the spectra kernel
#endif
template <int N2>
__global__ void __launch_bounds__(32 * RTB_SPEC_NW)
sh_spectra_soa_kernel(const ShPartParams prm, float2 *__restrict__ zspecT, long long Spad) {
    using Cfg = FftCfg<N2>;
    constexpr int G = Cfg::G, EPL = Cfg::EPL, GROUPS = Cfg::GROUPS, B = N2 / 2;
    constexpr int NW = RTB_SPEC_NW, ROWS = GROUPS * N2, RS = NW + 1;   // tile rows = (signal of the round, bin), padded
    __shared__ float2 s_tw[N2];
    __shared__ __align__(128) float2 s_buf[NW][GROUPS * Cfg::REGION];
    extern __shared__ __align__(16) float2 s_t[];   // [ROWS][RS] transposition tile (dynamic: > 48 KB in total)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane / G, l = lane % G;
    for (int i = threadIdx.x; i < N2; i += 32 * NW) s_tw[i] = prm.tw[i];
    __syncthreads();
    const int s0 = blockIdx.x * NW, s = s0 + warp;
    const bool active = s < prm.S;
    float2 *buf = s_buf[warp] + grp * Cfg::REGION;
    float *zrow = reinterpret_cast<float *>(buf);
    // blockIdx.y takes a contiguous share of the signal rounds: with one CTA walking all rounds the grid
    // of c3 (512 CTAs of 8 streams, 444 resident) ran as one full wave and a tail of 68 CTAs
    const int np_all = max(prm.np_cur, prm.np_last);
    const int share = ((np_all + (int)gridDim.y - 1) / (int)gridDim.y + GROUPS - 1) / GROUPS * GROUPS;
    const int p_first = (int)blockIdx.y * share, np = min(np_all, p_first + share);
    const size_t zoff = (size_t)(active ? s : 0) * prm.J * B;

    auto prefetch_z = [&](int p0) {
        const int p = p0 + grp;
        if (p >= np || !active) return;
        const ShSignal sg = prm.signals[p];
        constexpr int CH = B / 4;
        for (int i = l; i < 4 * CH; i += G) {
            const int part = i / CH, off = (i % CH) * 4;
            const int row = (part < 2) ? sg.row_re : sg.row_im;
            float *dst = zrow + (part >> 1) * N2 + (part & 1) * B + off;
            if (row >= 0)
                cp_async16_cg(dst, ((part & 1) ? prm.zcur : prm.zprev) + zoff + (size_t)row * B + off);
            else
                *reinterpret_cast<float4 *>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    if (np > p_first) prefetch_z(p_first);
    cp_async_commit();
    for (int p0 = p_first; p0 < np; p0 += GROUPS) {
        cp_async_wait_all();
        __syncwarp();
        const int p = p0 + grp;
        const bool valid = p < np && active;
        float2 X[EPL];
#pragma unroll
        for (int u = 0; u < EPL; ++u) {
            X[u].x = valid ? zrow[l + G * u] : 0.f;
            X[u].y = valid ? zrow[N2 + l + G * u] : 0.f;
        }
        __syncwarp();
        warp_fft_hook<N2, false>(X, buf, s_tw, l, [&] {
            if (p0 + GROUPS < np) prefetch_z(p0 + GROUPS);
            cp_async_commit();
        });
        __syncthreads();  // the previous round's tile has been written out
#pragma unroll
        for (int u = 0; u < EPL; ++u) s_t[(grp * N2 + l + G * u) * RS + warp] = X[u];
        __syncthreads();
        for (int i = threadIdx.x; i < ROWS * NW; i += 32 * NW) {
            const int row = i / NW, w = i % NW;
            const int pg = row / N2, bin = row % N2;
            if (p0 + pg < np && s0 + w < prm.S)
                zspecT[((size_t)bin * prm.NPT + (p0 + pg)) * Spad + s0 + w] = s_t[row * RS + w];
        }
    }
    cp_async_wait_all();
}

// Queue heads -> inverse FFTs -> cross-fade, queues in the stream-minor layout.
template <int N2>
__global__ void __launch_bounds__(128)
sh_partition_tail_soa_kernel(const ShPartParams prm, float2 *__restrict__ qcurT, float2 *__restrict__ qlastT,
                             long long Spad, int clear_heads) {
    using Cfg = FftCfg<N2>;
    constexpr int G = Cfg::G, EPL = Cfg::EPL, HB = EPL / 2, GROUPS = Cfg::GROUPS, B = N2 / 2;
    __shared__ float2 s_tw[N2];
    __shared__ __align__(128) float2 s_buf[4][GROUPS * Cfg::REGION];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane / G, l = lane % G;
    for (int i = threadIdx.x; i < N2; i += 128) s_tw[i] = prm.tw[i];
    __syncthreads();
    const int s = blockIdx.x * 4 + warp;
    if (s >= prm.S) return;
    float2 acc[HB + 1][4];
    const size_t fstride = (size_t)prm.P * 2 * Spad;
    float2 *qc = qcurT + (size_t)prm.head_cur * 2 * Spad + s;    // + f * fstride + ear * Spad
    float2 *ql = qlastT + (size_t)prm.head_last * 2 * Spad + s;
    const float2 zero = make_float2(0.f, 0.f);
#pragma unroll
    for (int u = 0; u <= HB; ++u) {
        const int f = (u < HB) ? (l + G * u) : B;
        acc[u][0] = qc[f * fstride];
        acc[u][1] = qc[f * fstride + Spad];
        acc[u][2] = prm.crossfade ? ql[f * fstride] : zero;
        acc[u][3] = prm.crossfade ? ql[f * fstride + Spad] : zero;
    }
    __syncwarp();  // every group has read the heads before anyone clears them
    // queue slot 0 leaves the queue: blocks[-1] = 0 after the roll.  With clear_heads == 0 the slot is
    // left as it is and the next block's partition MAC overwrites it (the plan keeps track).
#pragma unroll
    for (int u = 0; u <= HB; ++u) {
        const int f = (u < HB) ? (l + G * u) : B;
        if (clear_heads && grp == 0 && (u < HB || l == 0)) {
            qc[f * fstride] = zero;
            qc[f * fstride + Spad] = zero;
            if (prm.crossfade) { ql[f * fstride] = zero; ql[f * fstride + Spad] = zero; }
        }
    }
    if (prm.crossfade && lane == 0)
        prm.rad_last_out[s] = yaw_deg_to_rad16(prm.yaw_deg[s], prm.tracker_dir, prm.src_azim16);
    render_tail<N2>(acc, s_buf[warp] + grp * Cfg::REGION, s_tw, grp, l, prm.win_in, prm.win_out,
                    prm.crossfade, prm.out + (size_t)s * 2 * B);
}

// passthrough with partitions, stream-minor queues: the head of the current queue is dropped
static __global__ void sh_queue_clear_head_soa_kernel(float2 *qT, int S, int F, int P, long long Spad, int head) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per = 2LL * S;
    if (i >= (long long)F * per) return;
    const long long f = i / per, r = i % per;
    const long long e = r / S, s = r % S;
    qT[(((size_t)f * P + head) * 2 + e) * Spad + s] = make_float2(0.f, 0.f);
}

}  // namespace rtb
