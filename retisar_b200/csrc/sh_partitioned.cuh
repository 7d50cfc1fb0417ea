// Row a-X (EXTENSION, not in the reference): SH-domain rendering with a P-partition HRIR.
//
// Semantics by composition of the reference's own partition-queue convention
// (AdjustableFdConvolver.filter_block, ReTiSAR/_convolver.py:802-836, queue roll :645-665): the
// current rotation is applied at input time into an OUTPUT-side queue of P ear spectra, a second
// queue receives the previous rotation, both heads are inverse-transformed and cross-faded, then
// the queues advance.  (An input-side delay line is not possible here: the rotation is time
// varying and the rotated SH spectra are K/E times larger than the ear spectra.)
//
// Three kernels per block, after the analysis GEMM:
//   sh_spectra_kernel        one warp per stream: FFT of every complex SH signal -> zspec (HBM)
//   sh_partition_mac_kernel  per (32 bins, 8 streams): for all partitions q and signal terms j
//                            queue[q][ear][f] += H[q][j][ear][f] * phase_j * Z_j(f)
//                            -- per bin a [P*E x 2NP] x [2NP x streams] complex GEMM whose matrix is
//                            shared by all streams; register tile 4 partitions x 2 ears x cur/last
//                            x 2 streams per lane, H rows reused from L1 by the CTA's warps
//   sh_partition_tail_kernel one warp per stream: queue heads -> inverse FFTs -> cross-fade
#pragma once

namespace rtb {

// one (signal, term) step of the contraction: which spectrum, mirrored or not, which phase
struct ShStep {
    int p;      // signal index
    int term;   // 0: bin f, 1: conj(bin N2-f)
    int m;      // phase exp(-i*m*alpha)
    int pad;
};

struct ShPartParams {
    const float *zprev, *zcur;   // [S][J][B] analysis rows, previous / current block
    float2 *zspec;               // [S][NPT][N2] spectra of the complex signals (scratch)
    const float4 *htab;          // [P][NPT][2][Fp] (ear0.re, ear0.im, ear1.re, ear1.im)
    const ShSignal *signals;     // [NPT]
    const struct ShStep *steps;  // flattened (signal, term) list, ordered by signal
    const float2 *tw;
    const float *win_in, *win_out, *yaw_deg, *rad_last_in;
    float *rad_last_out;
    float2 *qcur, *qlast;        // [S][P][2][Fp] output-side queues (rings over P)
    float *out;                  // [S][2][B]
    int S, J, P, Fp, NPT;
    int np_cur, np_last, order_max, crossfade, head_cur, head_last;
    int nsteps_cur, nsteps_last;  // steps (prefix) taking part in the current / previous yaw sum
    float tracker_dir, src_azim16;
};

template <int N2>
__global__ void __launch_bounds__(128)
sh_spectra_kernel(const ShPartParams prm) {
    using Cfg = FftCfg<N2>;
    constexpr int G = Cfg::G, EPL = Cfg::EPL, GROUPS = Cfg::GROUPS, B = N2 / 2;
    __shared__ float2 s_tw[N2];
    __shared__ __align__(128) float2 s_buf[4][GROUPS * Cfg::REGION];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane / G, l = lane % G;
    for (int i = threadIdx.x; i < N2; i += 128) s_tw[i] = prm.tw[i];
    __syncthreads();
    const int s = blockIdx.x * 4 + warp;
    if (s >= prm.S) return;
    float2 *buf = s_buf[warp] + grp * Cfg::REGION;
    float *zrow = reinterpret_cast<float *>(buf);
    const int np = max(prm.np_cur, prm.np_last);
    const size_t zoff = (size_t)s * prm.J * B;

    auto prefetch_z = [&](int p0) {
        const int p = p0 + grp;
        if (p >= np) return;
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
    if (np > 0) prefetch_z(0);
    cp_async_commit();
    for (int p0 = 0; p0 < np; p0 += GROUPS) {
        cp_async_wait_all();
        __syncwarp();
        const int p = p0 + grp;
        const bool valid = p < np;
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
        if (valid) {
            float2 *dst = prm.zspec + ((size_t)s * prm.NPT + p) * N2 + l;
#pragma unroll
            for (int u = 0; u < EPL; ++u) dst[G * u] = X[u];
        }
    }
    cp_async_wait_all();
}

#define RTB_QC 4  // partitions per register tile

// The partition MAC (SIMT form): the H rows of the CTA's 32 bins are staged through shared
// memory (cp.async, double buffered over chunks of RTB_MAC_SC steps) and shared by 8 warps = 16
// streams, instead of every warp pulling them through L1/L2 (first version: 5.1 long-scoreboard
// stall cycles per issue, FMA pipe 33 %, profiles/).  Spectra are read in batches (see the loop).
#define RTB_MAC_SC 8        // steps per staged chunk
#ifndef RTB_MAC_ZB
#define RTB_MAC_ZB 2        // steps per batch of spectrum loads (measured c3: 2 -> 2479 us, 4 -> 2503 us, per-step prefetch 2639 us)
#endif
#define RTB_MAC_WARPS 8     // warps per CTA, 2 streams each
#define RTB_MAC_MAX_STEPS 192

template <int N2>
__global__ void __launch_bounds__(32 * RTB_MAC_WARPS, 2)
sh_partition_mac2_kernel(const ShPartParams prm) {
    constexpr int B = N2 / 2;
    constexpr int NT = 32 * RTB_MAC_WARPS;
    __shared__ __align__(16) float4 s_h[2][RTB_MAC_SC][RTB_QC][32];
    __shared__ float2 s_cs[RTB_MAC_WARPS][2][2][RTB_MAX_ORDER + 1];  // [warp][stream][cur/last][m]
    __shared__ ShStep s_steps[RTB_MAC_MAX_STEPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f0 = blockIdx.x * 32;
    const int f = f0 + lane;
    const bool fvalid = f <= B;
    const int fc = fvalid ? f : 0;                 // clamped: out-of-range lanes read bin 0
    const int fm = (N2 - fc) & (N2 - 1);           // mirrored bin
    const int sbase = (blockIdx.y * RTB_MAC_WARPS + warp) * 2;
    const bool wactive = sbase < prm.S;
    const bool has_s1 = sbase + 1 < prm.S;
    const int nsteps = max(prm.nsteps_cur, prm.nsteps_last);
    for (int i = threadIdx.x; i < nsteps; i += NT) s_steps[i] = prm.steps[i];
    if (wactive) {
        for (int i = lane; i < 4 * (prm.order_max + 1); i += 32) {
            const int m = i % (prm.order_max + 1), which = i / (prm.order_max + 1);
            const int st = which >> 1, cl = which & 1;
            const int s = min(sbase + st, prm.S - 1);
            const float rad = cl ? prm.rad_last_in[s]
                                 : yaw_deg_to_rad16(prm.yaw_deg[s], prm.tracker_dir, prm.src_azim16);
            float sn, cs;
            sincosf((float)m * rad, &sn, &cs);
            s_cs[warp][st][cl][m] = make_float2(cs, sn);
        }
    }
    __syncthreads();
    const float2 *z0 = prm.zspec + (size_t)(wactive ? sbase : 0) * prm.NPT * N2;
    const float2 *z1 = prm.zspec + (size_t)(has_s1 ? sbase + 1 : (wactive ? sbase : 0)) * prm.NPT * N2;
    const size_t hq = (size_t)prm.NPT * 2 * prm.Fp;  // H stride between partitions
    const int nchunks = (nsteps + RTB_MAC_SC - 1) / RTB_MAC_SC;
    // the last bin block holds the Nyquist bin only: clamp the staged columns into the row
    const int fcol = min(f0 + lane, prm.Fp - 1);

    for (int q0 = 0; q0 < prm.P; q0 += RTB_QC) {
        // stage chunk c of this partition group: rows (step, partition) x 32 bins
        auto stage_h = [&](int c, int buf) {
            for (int i = threadIdx.x; i < RTB_MAC_SC * RTB_QC * 32; i += NT) {
                const int col = i & 31, a = (i >> 5) % RTB_QC, jj = i / (32 * RTB_QC);
                const int j = c * RTB_MAC_SC + jj;
                float4 *dst = &s_h[buf][jj][a][col];
                if (j < nsteps && q0 + a < prm.P) {
                    const ShStep st_ = s_steps[j];
                    const float4 *src = prm.htab + (size_t)(q0 + a) * hq + ((size_t)st_.p * 2 + st_.term) * prm.Fp
                                        + min(f0 + col, prm.Fp - 1);
                    cp_async16_cg(dst, src);
                } else {
                    *dst = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        };
        (void)fcol;
        float2 acc[RTB_QC][2][2][2];  // [partition][ear][cur/last][stream]
#pragma unroll
        for (int a = 0; a < RTB_QC; ++a)
#pragma unroll
            for (int e = 0; e < 2; ++e)
#pragma unroll
                for (int c = 0; c < 2; ++c) acc[a][e][c][0] = acc[a][e][c][1] = make_float2(0.f, 0.f);
        if (nchunks > 0) stage_h(0, 0);
        cp_async_commit();
        // Spectra: loaded in batches of RTB_MAC_ZB steps into two register rings.  ptxas tracks every
        // LDG of this loop on one scoreboard, so the first use of ANY loaded value waits for ALL
        // loads in flight: a per-step prefetch ("two steps ahead") in fact waited for the load issued
        // in the same step (long-scoreboard 4.3 stall cycles per issue, FMA pipe 40 %).  With a batch
        // issued every RTB_MAC_ZB steps the wait comes RTB_MAC_ZB steps of arithmetic later.
        static_assert(RTB_MAC_SC % (2 * RTB_MAC_ZB) == 0, "chunk = whole pairs of spectrum batches");
        float2 zr[2][RTB_MAC_ZB][2];  // [ring][step in batch][stream]
        auto load_z = [&](int j, float2 (&z)[2]) {
            if (j < nsteps && wactive) {
                const ShStep st_ = s_steps[j];
                const size_t zo = (size_t)st_.p * N2 + (st_.term ? fm : fc);
                z[0] = __ldg(z0 + zo);
                z[1] = __ldg(z1 + zo);
            } else {
                z[0] = z[1] = make_float2(0.f, 0.f);
            }
        };
#pragma unroll
        for (int i = 0; i < RTB_MAC_ZB; ++i) load_z(i, zr[0][i]);
        for (int c = 0; c < nchunks; ++c) {
            cp_async_wait_all();
            __syncthreads();  // chunk c has landed; everyone has left chunk c-1 (its buffer is free)
            if (c + 1 < nchunks) stage_h(c + 1, (c + 1) & 1);
            cp_async_commit();
            if (!wactive) continue;
#pragma unroll
            for (int jj = 0; jj < RTB_MAC_SC; ++jj) {
                const int j = c * RTB_MAC_SC + jj;
                constexpr int ZB = RTB_MAC_ZB;
                const int ring = (jj / ZB) & 1, slot = jj % ZB;
                const float2 zz[2] = {zr[ring][slot][0], zr[ring][slot][1]};
                if (slot == 0) {  // first use of this batch (the wait): now request the next one
#pragma unroll
                    for (int i = 0; i < ZB; ++i) load_z(j + ZB + i, zr[ring ^ 1][i]);
                }
                if (j < nsteps) {
                    const ShStep stp = s_steps[j];
                    const bool in_cur = j < prm.nsteps_cur, in_last = j < prm.nsteps_last;
                    float2 A[2][2];  // [cur/last][stream]
#pragma unroll
                    for (int st = 0; st < 2; ++st) {
                        float2 z = zz[st];
                        if (stp.term) z.y = -z.y;
                        const float2 pc = in_cur ? phase_of(s_cs[warp][st][0], stp.m) : make_float2(0.f, 0.f);
                        const float2 pl = in_last ? phase_of(s_cs[warp][st][1], stp.m) : make_float2(0.f, 0.f);
                        A[0][st] = cmul(pc, z);
                        A[1][st] = cmul(pl, z);
                    }
#pragma unroll
                    for (int a = 0; a < RTB_QC; ++a) {
                        const float4 h = s_h[c & 1][jj][a][lane];
#pragma unroll
                        for (int cl = 0; cl < 2; ++cl)
#pragma unroll
                            for (int st = 0; st < 2; ++st) {
                                cfma(acc[a][0][cl][st], make_float2(h.x, h.y), A[cl][st]);
                                cfma(acc[a][1][cl][st], make_float2(h.z, h.w), A[cl][st]);
                            }
                    }
                }
            }
        }
        __syncthreads();  // the next partition group restages buffer 0
        if (fvalid && wactive) {
#pragma unroll
            for (int a = 0; a < RTB_QC; ++a) {
                if (q0 + a < prm.P) {
                    const int sc = (prm.head_cur + q0 + a) % prm.P, sl = (prm.head_last + q0 + a) % prm.P;
#pragma unroll
                    for (int st = 0; st < 2; ++st) {
                        if (st == 0 || has_s1) {
                            const size_t sq = (size_t)(sbase + st) * prm.P;
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                float2 *qc = prm.qcur + ((sq + sc) * 2 + e) * prm.Fp + f;
                                float2 v = *qc;
                                v.x += acc[a][e][0][st].x; v.y += acc[a][e][0][st].y;
                                *qc = v;
                                if (prm.crossfade) {
                                    float2 *ql = prm.qlast + ((sq + sl) * 2 + e) * prm.Fp + f;
                                    float2 w = *ql;
                                    w.x += acc[a][e][1][st].x; w.y += acc[a][e][1][st].y;
                                    *ql = w;
                                }
                            }
                        }
                    }
                }
            }
        }
    }
}

template <int N2>
__global__ void __launch_bounds__(128)
sh_partition_tail_kernel(const ShPartParams prm) {
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
    float2 *qc = prm.qcur + ((size_t)s * prm.P + prm.head_cur) * 2 * prm.Fp;
    float2 *ql = prm.qlast + ((size_t)s * prm.P + prm.head_last) * 2 * prm.Fp;
    const float2 zero = make_float2(0.f, 0.f);
#pragma unroll
    for (int u = 0; u <= HB; ++u) {
        const int f = (u < HB) ? (l + G * u) : B;
        acc[u][0] = qc[f];
        acc[u][1] = qc[prm.Fp + f];
        acc[u][2] = prm.crossfade ? ql[f] : zero;
        acc[u][3] = prm.crossfade ? ql[prm.Fp + f] : zero;
    }
    __syncwarp();  // every group has read the heads before anyone clears them
#pragma unroll
    for (int u = 0; u <= HB; ++u) {  // queue slot 0 leaves the queue: blocks[-1] = 0 after the roll
        const int f = (u < HB) ? (l + G * u) : B;
        if (grp == 0 && (u < HB || l == 0)) {
            qc[f] = zero;
            qc[prm.Fp + f] = zero;
            if (prm.crossfade) { ql[f] = zero; ql[prm.Fp + f] = zero; }
        }
    }
    if (prm.crossfade && lane == 0)
        prm.rad_last_out[s] = yaw_deg_to_rad16(prm.yaw_deg[s], prm.tracker_dir, prm.src_azim16);
    render_tail<N2>(acc, s_buf[warp] + grp * Cfg::REGION, s_tw, grp, l, prm.win_in, prm.win_out,
                    prm.crossfade, prm.out + (size_t)s * 2 * B);
}

// passthrough with partitions: the head of the current queue is dropped and the queue advances
static __global__ void sh_queue_clear_head_kernel(float2 *q, int S, int P, int Fp, int head) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per = 2LL * Fp;
    if (i >= (long long)S * per) return;
    const long long s = i / per, r = i % per;
    q[((size_t)s * P + head) * per + r] = make_float2(0.f, 0.f);
}

}  // namespace rtb
