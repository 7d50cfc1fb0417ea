// Uniformly partitioned overlap-save FIR convolver for S batched streams (reference:
// OverlapSaveConvolver, ReTiSAR/_convolver.py:436-669; used for the ARIR pre-renderer, the
// headphone compensation filter and the reference's own benchmark).
//
// Formulation.  The reference keeps an OUTPUT-side queue: every block adds H[p] * X_t into queue
// slot p for all partitions p and emits slot 0 (:602-622, :645-665).  The filters are time
// invariant, so the same output is y_t = irfft( sum_p H[p] * X_(t-p) ): an INPUT-side delay line of
// the last P input spectra.  That needs one spectrum write plus P spectrum reads per block instead
// of P read-modify-writes, i.e. half the HBM traffic for 2-in/2-out and ~Cout times less for the
// mono-in ARIR case, and it is what this kernel implements (the "FDL kernel" of the north star).
// Passthrough (:558-562) overrides queue slot 0 with X_t and adds nothing to later slots; here the
// block is flagged so that later sums skip it.
//
// One warp owns a pair of output channels of one stream: a single complex FFT transforms both
// real overlap-save buffers (or the mono buffer), the multiply-accumulate over the delay line
// runs in registers on the bins the lane owns, and one complex inverse FFT returns both channels.
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <new>
#include <vector>

#include "fft_warp.cuh"
#include "sh_analysis_tc.cuh"

using namespace rtb;

namespace {

struct OlsParams {
    const float *x;            // [S][Cin][B] current input block
    const float *xprev;        // [S][Cin][B] previous input block (state, read side)
    float *xnext;              // [S][Cin][B] state, write side (ping-pong)
    float2 *hist;              // [S][P][Cin][Fp] delay line of input spectra, ring over P
    const float2 *hfd;         // [P][Cout][Fp] filter partition spectra
    unsigned char *incl;       // [P] ring slot takes part in later sums (0 after a passthrough block)
    const float2 *tw;          // [N2]
    float *out;                // [S][Cout][B]
    int S, Cin, Cout, P, Fp, slot, passthrough;
    int TS, TP, tiles_s;       // CTA tile: TS streams x TP channel pairs; tiles along the stream axis
};

template <int N2>
__global__ void __launch_bounds__(FftCfg<N2>::THREADS)
ols_fdl_kernel(const OlsParams prm) {
    using Cfg = FftCfg<N2>;
    constexpr int G = Cfg::G, EPL = Cfg::EPL, HB = EPL / 2, GROUPS = Cfg::GROUPS;
    constexpr int B = N2 / 2;
    constexpr bool BIG = Cfg::BIG;  // one transform per CTA, exchange buffer in dynamic smem

    extern __shared__ __align__(128) unsigned char dyn_smem[];
    __shared__ float2 s_tw[BIG ? 1 : N2];
    __shared__ __align__(128) float2 s_buf[BIG ? 1 : 4 * GROUPS * Cfg::REGION];

    const int warp = BIG ? 0 : (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int grp = BIG ? 0 : lane / G, l = BIG ? (int)threadIdx.x : lane % G;
    const float2 *tw = prm.tw;
    float2 *buf = reinterpret_cast<float2 *>(dyn_smem);
    if constexpr (!BIG) {
        for (int i = threadIdx.x; i < N2; i += 128) s_tw[i] = prm.tw[i];
        __syncthreads();
        tw = s_tw;
        buf = s_buf + (warp * GROUPS + grp) * Cfg::REGION;
    }

    // A CTA covers a tile of TS streams x TP channel pairs (TS * TP = transforms per CTA): the
    // groups of a warp are different streams of the same pair, so their filter loads coalesce, and
    // the warps of a CTA share the streams' delay lines through L1 (mono-in ARIR case: the filter
    // and the delay line are each re-read from L2 1 / TS and 1 / TP as often as with one task per
    // (stream, pair) in stream order; measured 1.46 -> ... ms for 1 024 streams x 434 channels x 64 partitions).
    const int npairs = (prm.Cout + 1) / 2;
    const int local = BIG ? 0 : warp * GROUPS + grp;
    const int tile_s = (int)(blockIdx.x % (unsigned)prm.tiles_s), tile_p = (int)(blockIdx.x / (unsigned)prm.tiles_s);
    const int s_raw = tile_s * prm.TS + local % prm.TS, pair_raw = tile_p * prm.TP + local / prm.TS;
    const bool valid = s_raw < prm.S && pair_raw < npairs;
    // a warp whose groups are all out of range has nothing to do (uniform exit)
    if constexpr (!BIG) {
        if (!__any_sync(0xffffffffu, valid)) return;
    } else {
        if (!valid) return;
    }
    const int s = valid ? s_raw : 0;
    const int pair = valid ? pair_raw : 0;
    const int co0 = 2 * pair, co1 = 2 * pair + 1;
    const bool has1 = co1 < prm.Cout;
    const bool mono = prm.Cin == 1 && prm.Cout > 1;
    const int ci0 = mono ? 0 : co0, ci1 = mono ? 0 : co1;

    // ---- a1: overlap-save buffers (previous | current block) of both channels, one complex FFT
    float2 X[EPL];
    {
        const size_t r0 = ((size_t)s * prm.Cin + ci0) * B + l;
        const size_t r1 = ((size_t)s * prm.Cin + ci1) * B + l;
        const bool second = valid && has1 && !mono;
#pragma unroll
        for (int u = 0; u < HB; ++u) {
            X[u].x = valid ? prm.xprev[r0 + G * u] : 0.f;
            X[u + HB].x = valid ? prm.x[r0 + G * u] : 0.f;
            X[u].y = second ? prm.xprev[r1 + G * u] : 0.f;
            X[u + HB].y = second ? prm.x[r1 + G * u] : 0.f;
        }
        // state: the current block becomes the previous one (written once per input channel)
        if (valid && (!mono || pair == 0)) {
#pragma unroll
            for (int u = 0; u < HB; ++u) {
                prm.xnext[r0 + G * u] = X[u + HB].x;
                if (second) prm.xnext[r1 + G * u] = X[u + HB].y;
            }
        }
    }
    warp_fft<N2, false>(X, buf, tw, l);

    // separate the two real signals on the bins this lane owns: slot u <-> bin l + G*u, slot HB
    // is the Nyquist bin (lane 0)
    float2 X0[HB + 1], X1[HB + 1];
    float2 M[HB + 1];
    fft_mirror<N2>(X, M, buf, l);
#pragma unroll
    for (int u = 0; u <= HB; ++u) {
        const float2 zf = X[u], zp = M[u];
        // Z = A + iB with A, B Hermitian: A(f) = (Z(f) + conj Z(-f)) / 2, B(f) = (Z(f) - conj Z(-f)) / 2i
        X0[u] = make_float2(0.5f * (zf.x + zp.x), 0.5f * (zf.y - zp.y));
        X1[u] = mono ? X0[u] : make_float2(0.5f * (zf.y + zp.y), 0.5f * (zp.x - zf.x));
    }

    // delay line: store X_t, then accumulate sum_p H[p] * X_(t-p)
    const int Fp = prm.Fp;
    const bool own_nyq = (l == 0);
    if (valid && (!mono || pair == 0)) {
        float2 *h0 = prm.hist + (((size_t)s * prm.P + prm.slot) * prm.Cin + ci0) * Fp;
        float2 *h1 = prm.hist + (((size_t)s * prm.P + prm.slot) * prm.Cin + ci1) * Fp;
#pragma unroll
        for (int u = 0; u <= HB; ++u) {
            const int f = l + G * u;
            if (u < HB || own_nyq) {
                h0[f] = X0[u];
                if (has1 && !mono) h1[f] = X1[u];
            }
        }
    }
    float2 Y0[HB + 1], Y1[HB + 1];
    if (prm.passthrough) {  // queue slot 0 := X_t (_convolver.py:558-562)
#pragma unroll
        for (int u = 0; u <= HB; ++u) { Y0[u] = X0[u]; Y1[u] = X1[u]; }
    } else {
        const float2 *g0 = prm.hfd + (size_t)co0 * Fp;
        const float2 *g1 = prm.hfd + (size_t)(has1 ? co1 : co0) * Fp;
#pragma unroll
        for (int u = 0; u <= HB; ++u) {
            const int f = l + G * u;
            const bool ok = valid && (u < HB || own_nyq);
            const float2 a = ok ? __ldg(g0 + f) : make_float2(0.f, 0.f);
            const float2 b = ok ? __ldg(g1 + f) : make_float2(0.f, 0.f);
            Y0[u] = cmul(a, X0[u]);
            Y1[u] = cmul(b, X1[u]);
        }
        int hs = prm.slot;
        for (int p = 1; p < prm.P; ++p) {
            hs = (hs == 0) ? prm.P - 1 : hs - 1;
            if (!prm.incl[hs]) continue;  // uniform: that block ran in passthrough
            const float2 *x0 = prm.hist + (((size_t)s * prm.P + hs) * prm.Cin + ci0) * Fp;
            const float2 *x1 = prm.hist + (((size_t)s * prm.P + hs) * prm.Cin + ci1) * Fp;
            const float2 *gp0 = g0 + (size_t)p * prm.Cout * Fp;
            const float2 *gp1 = g1 + (size_t)p * prm.Cout * Fp;
            float2 xa[HB + 1], xb[HB + 1], ga[HB + 1], gb[HB + 1];
#pragma unroll
            for (int u = 0; u <= HB; ++u) {
                const int f = l + G * u;
                const bool ok = valid && (u < HB || own_nyq);
                xa[u] = ok ? __ldcs(x0 + f) : make_float2(0.f, 0.f);
                ga[u] = ok ? __ldg(gp0 + f) : make_float2(0.f, 0.f);
                if (!mono) xb[u] = (ok && has1) ? __ldcs(x1 + f) : make_float2(0.f, 0.f);
                gb[u] = (ok && has1) ? __ldg(gp1 + f) : make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u <= HB; ++u) {
                cfma(Y0[u], ga[u], xa[u]);
                cfma(Y1[u], gb[u], mono ? xa[u] : xb[u]);
            }
        }
    }

    // ---- a6: Hermitian extension of (Y0 + i Y1), inverse FFT, keep the second half
    {
        float2 V[HB];
#pragma unroll
        for (int u = 0; u < HB; ++u) {
            float2 eL = Y0[u], eR = has1 ? Y1[u] : make_float2(0.f, 0.f);
            if (u == 0 && l == 0) { eL.y = 0.f; eR.y = 0.f; }
            X[u] = make_float2(eL.x - eR.y, eL.y + eR.x);
            V[u] = make_float2(eL.x + eR.y, eR.x - eL.y);
        }
        const float2 nyq = make_float2(Y0[HB].x, has1 ? Y1[HB].x : 0.f);
        fft_hermitian_fill<N2>(X, V, nyq, buf, l);
    }
    fft_sync<N2>();
    warp_fft<N2, true>(X, buf, tw, l);
    if (valid) {
        const float sc = 1.0f / (float)N2;
        float *o0 = prm.out + ((size_t)s * prm.Cout + co0) * B + l;
        float *o1 = prm.out + ((size_t)s * prm.Cout + (has1 ? co1 : co0)) * B + l;
#pragma unroll
        for (int u = 0; u < HB; ++u) {
            o0[G * u] = X[u + HB].x * sc;
            if (has1) o1[G * u] = X[u + HB].y * sc;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) prm.incl[prm.slot] = prm.passthrough ? 0 : 1;
}


// ---------------------------------------------------------------------------------------------
// Mono input, many outputs (the ARIR pre-renderer: 1 -> 434 channels, 64 partitions at 64-sample
// blocks).  With one warp group per (stream, channel pair) the kernel above re-reads every filter
// partition per stream and runs at the load-issue rate (15 loads for 20 packed FMAs per partition:
// 1.46 ms per block of 1 024 streams, 1 % of the HBM roofline).  Batched over streams the sum over
// partitions is a GEMM per frequency bin f,
//     Y[f][c][s] = sum_p H[p][c][f] * X[s][t-p][f]      (complex; M = streams, K = partitions, N = channels)
// so it runs on the tcgen05 kernel of the SH analysis (sh_analysis_tc.cuh) with the bin in the role
// of the "stream" (one operand matrix per bin, streamed through the stages), the streams in the
// role of the samples and the complex product embedded in real arithmetic:
//     [Yr Yi] = [Xr Xi] [[Hr, Hi], [-Hi, Hr]],  K index = (age p, re / im), N index = (channel, re / im).
//   ols_mono_spectrum_kernel  FFT of the mono overlap-save buffer -> delay line [f][k][s] (age 0 rows)
//   (cudaMemcpy2D)            ages the delay line by one partition (ping-pong buffers)
//   sh_analysis_tc_kernel     one launch per chunk of 112 channels (N <= 256 TMEM columns)
//   ols_mono_ifft_kernel      gathers the bins of a channel pair for 32 streams through shared
//                             memory, Hermitian extension, inverse FFT, overlap-save discard
// ---------------------------------------------------------------------------------------------
struct OlsMonoParams {
    const float *x, *xprev;    // [S][B]
    float *xnext;
    float *histT;              // [F][Kpad][Spad]: rows 2*slot, 2*slot + 1 of every bin = Re / Im of X_t
    int slot;                  // ring slot of this block (the delay line ages by moving the slot, not the data)
    const float2 *tw;
    int S, Spad, Kpad, passthrough;
};

template <int N2>
__global__ void __launch_bounds__(128)
ols_mono_spectrum_kernel(const OlsMonoParams prm) {
    using Cfg = FftCfg<N2>;
    constexpr int G = Cfg::G, EPL = Cfg::EPL, HB = EPL / 2, GROUPS = Cfg::GROUPS, B = N2 / 2;
    __shared__ float2 s_tw[N2];
    __shared__ __align__(128) float2 s_buf[4 * GROUPS * Cfg::REGION];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane / G, l = lane % G;
    for (int i = threadIdx.x; i < N2; i += 128) s_tw[i] = prm.tw[i];
    __syncthreads();
    float2 *buf = s_buf + (warp * GROUPS + grp) * Cfg::REGION;
    const int s_raw = (blockIdx.x * 4 + warp) * GROUPS + grp;
    const bool valid = s_raw < prm.S;
    if (!__any_sync(0xffffffffu, valid)) return;
    const int s = valid ? s_raw : 0;
    float2 X[EPL];
    const size_t r0 = (size_t)s * B + l;
#pragma unroll
    for (int u = 0; u < HB; ++u) {
        X[u] = make_float2(valid ? prm.xprev[r0 + G * u] : 0.f, 0.f);
        X[u + HB] = make_float2(valid ? prm.x[r0 + G * u] : 0.f, 0.f);
    }
    if (valid) {
#pragma unroll
        for (int u = 0; u < HB; ++u) prm.xnext[r0 + G * u] = X[u + HB].x;
    }
    warp_fft<N2, false>(X, buf, s_tw, l);
    if (valid) {
        // a passthrough block takes no part in later sums (_convolver.py:558-562): store zeros
        const float keep = prm.passthrough ? 0.f : 1.f;
        const size_t fstride = (size_t)prm.Kpad * prm.Spad;
#pragma unroll
        for (int u = 0; u <= HB; ++u) {
            const int f = (u < HB) ? l + G * u : B;
            if (u < HB || l == 0) {
                // the input is real, so bin f of the transform is the spectrum itself; the Nyquist
                // bin sits in element HB of lane 0
                const float2 z = (u < HB) ? X[u] : X[HB];
                float *h = prm.histT + (size_t)f * fstride + (size_t)(2 * prm.slot) * prm.Spad + s;
                h[0] = z.x * keep;
                h[prm.Spad] = z.y * keep;
            }
        }
    }
}

struct OlsMonoIfftParams {
    const float *Y;            // chunk: [F][Jc][Spad], rows (2 * local channel, + 1) = Re, Im
    const float *x;            // [S][B] (passthrough: the input block is the output of every channel)
    const float2 *tw;
    float *out;                // [S][Cout][B]
    int S, Spad, Jc, Cout, c0, nch, passthrough;  // chunk covers channels c0 .. c0 + nch - 1
};

#ifndef RTB_OLS_MONO_ST
// streams per CTA of the gather kernel.  Measured (c4 chain with the folded pre-renderer, 1 024 / 4 096
// streams): 8 -> 0.2425 / 0.762 ms, 16 -> 0.2334 / 0.698 ms, 32 -> 0.2416 / 0.730 ms, 64 -> 0.2586 / 0.807 ms
#define RTB_OLS_MONO_ST 16
#endif
#ifndef RTB_OLS_MONO_LB
#define RTB_OLS_MONO_LB 5   // gather loads in flight per thread
#endif

template <int N2>
__global__ void __launch_bounds__(128)
ols_mono_ifft_kernel(const OlsMonoIfftParams prm) {
    using Cfg = FftCfg<N2>;
    constexpr int G = Cfg::G, EPL = Cfg::EPL, HB = EPL / 2, GROUPS = Cfg::GROUPS, B = N2 / 2, F = B + 1;
    constexpr int ST = RTB_OLS_MONO_ST, TASKS = 4 * GROUPS;
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    float *tile = reinterpret_cast<float *>(dyn_smem);  // [F][4 rows: re0 im0 re1 im1][ST + 1]
    __shared__ float2 s_tw[N2];
    __shared__ __align__(128) float2 s_buf[4 * GROUPS * Cfg::REGION];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane / G, l = lane % G;
    const int s0 = blockIdx.x * ST;
    const int pair = blockIdx.y;                 // channel pair inside the chunk
    const int cl0 = 2 * pair, cl1 = 2 * pair + 1;
    const bool has1 = cl1 < prm.nch;
    for (int i = threadIdx.x; i < N2; i += 128) s_tw[i] = prm.tw[i];
    // coalesced gather: for every bin the four rows of the pair, ST consecutive streams each, as
    // 16-byte loads (Spad is a multiple of ST); the scalar stores into the padded rows are
    // conflict free (bank = 33 r + 4 sl4 + j)
    // (all loads of a batch are requested before the first one is stored: ncu showed 15 long-scoreboard
    // stall cycles per issue with one load in flight per thread)
    constexpr int NLD = F * 4 * (ST / 4), PER = (NLD + 127) / 128;
    for (int k0 = 0; k0 < PER; k0 += RTB_OLS_MONO_LB) {
        float4 v[RTB_OLS_MONO_LB];
#pragma unroll
        for (int k = 0; k < RTB_OLS_MONO_LB; ++k) {
            const int i = threadIdx.x + 128 * (k0 + k);
            const int sl4 = i % (ST / 4), r = (i / (ST / 4)) & 3, f = i / ST;
            const int row = 2 * cl0 + r;  // rows of channel cl0 then cl1
            v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < NLD && (r < 2 || has1))
                v[k] = __ldcs(reinterpret_cast<const float4 *>(prm.Y + ((size_t)f * prm.Jc + row) * prm.Spad + s0) + sl4);
        }
#pragma unroll
        for (int k = 0; k < RTB_OLS_MONO_LB; ++k) {
            const int i = threadIdx.x + 128 * (k0 + k);
            const int sl4 = i % (ST / 4), r = (i / (ST / 4)) & 3, f = i / ST;
            if (i < NLD) {
                float *d = tile + (f * 4 + r) * (ST + 1) + 4 * sl4;
                d[0] = v[k].x; d[1] = v[k].y; d[2] = v[k].z; d[3] = v[k].w;
            }
        }
    }
    __syncthreads();
    float2 *buf = s_buf + (warp * GROUPS + grp) * Cfg::REGION;
    for (int base = 0; base < ST; base += TASKS) {
        const int sl = base + warp * GROUPS + grp;
        const int s = s0 + sl;
        const bool valid = sl < ST && s < prm.S;
        if (!__any_sync(0xffffffffu, valid)) continue;
        const int slc = valid ? sl : 0;
        float2 X[EPL], V[HB];
#pragma unroll
        for (int u = 0; u < HB; ++u) {
            const int f = l + G * u;
            float2 eL = make_float2(tile[(f * 4 + 0) * (ST + 1) + slc], tile[(f * 4 + 1) * (ST + 1) + slc]);
            float2 eR = make_float2(tile[(f * 4 + 2) * (ST + 1) + slc], tile[(f * 4 + 3) * (ST + 1) + slc]);
            if (u == 0 && l == 0) { eL.y = 0.f; eR.y = 0.f; }
            X[u] = make_float2(eL.x - eR.y, eL.y + eR.x);
            V[u] = make_float2(eL.x + eR.y, eR.x - eL.y);
        }
        const float2 nyq = make_float2(tile[(B * 4 + 0) * (ST + 1) + slc], tile[(B * 4 + 2) * (ST + 1) + slc]);
        fft_hermitian_fill<N2>(X, V, nyq, buf, l);
        fft_sync<N2>();
        warp_fft<N2, true>(X, buf, s_tw, l);
        if (valid) {
            const float sc = 1.0f / (float)N2;
            float *o0 = prm.out + ((size_t)s * prm.Cout + prm.c0 + cl0) * B + l;
#pragma unroll
            for (int u = 0; u < HB; ++u) {
                o0[G * u] = X[u + HB].x * sc;
                if (has1) o0[B + G * u] = X[u + HB].y * sc;
            }
        }
        __syncwarp();
    }
}

// passthrough of the mono convolver: every output channel carries the input block
__global__ void ols_mono_passthrough_kernel(const float *x, float *out, int S, int Cout, int B) {
    const size_t n = (size_t)S * Cout * B;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t s = i / ((size_t)Cout * B);
        out[i] = x[s * B + i % B];
    }
}

}  // namespace

struct rtb_ols_plan {
    int device = 0;
    int S = 0, B = 0, Cin = 0, Cout = 0, P = 0, F = 0, Fp = 0, N2 = 0;
    bool filter_set = false;
    float *d_xstate = nullptr;   // [2][S][Cin][B]
    float2 *d_hist = nullptr;    // [S][P][Cin][Fp]
    float2 *d_hfd = nullptr;     // [P][Cout][Fp]
    unsigned char *d_incl = nullptr;
    float2 *d_tw = nullptr;
    float *d_in = nullptr, *d_out = nullptr;  // staging of the host entry point
    cudaStream_t own_stream = nullptr;
    int slot = 0, par = 0;
    bool passthrough = false;
    long long launches = 0;
    bool profiling = false;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    bool ev_valid = false;
    // mono-in / many-out on the tensor cores (see ols_mono_* above)
    bool tc_mono = false;
    int tc_Spad = 0, tc_Kpad = 0, tc_nchunks = 0, tc_hpar = 0;
    int tc_nch[8] = {0}, tc_Jc[8] = {0}, tc_Npad[8] = {0}, tc_cols[8] = {0}, tc_smem[8] = {0};
    float *d_histT[2] = {nullptr, nullptr};   // [0]: [F][Kpad][Spad] delay line, a ring over the partition rows (tc_hpar = slot of the newest block)
    float *d_bmat[8] = {nullptr};             // per chunk: [F][hi, lo][Kpad/4][Npad/8][8][4]
    float *d_Y[8] = {nullptr};                // per chunk: [F][Jc][Spad]
};

namespace {
size_t hist_elems(const rtb_ols_plan *p) { return (size_t)p->S * p->P * p->Cin * p->Fp; }
size_t state_elems(const rtb_ols_plan *p) { return (size_t)p->S * p->Cin * p->B; }

// one block of the mono-in / many-out convolver on the tensor cores (see ols_mono_* above)
int process_mono_tc(rtb_ols_plan *p, const float *d_in, float *d_out, cudaStream_t st) {
    const int F = p->F, Kpad = p->tc_Kpad, Spad = p->tc_Spad;
    // The delay line ages by moving the ring slot: the newest spectrum goes to slot head, partition p
    // (age p) is read from slot (head + p) % P -- the GEMM's loaders rotate the K rows by 2 * head.
    // (Until round 2 the whole [F][Kpad][Spad] line was copied by one partition every block.)
    const int head = p->tc_hpar = (p->tc_hpar + p->P - 1) % p->P;
    float *nxt = p->d_histT[0];
    OlsMonoParams mp;
    mp.x = d_in;
    mp.xprev = p->d_xstate + (size_t)p->par * state_elems(p);
    mp.xnext = p->d_xstate + (size_t)(p->par ^ 1) * state_elems(p);
    mp.histT = nxt; mp.slot = head; mp.tw = p->d_tw; mp.S = p->S; mp.Spad = Spad; mp.Kpad = Kpad;
    mp.passthrough = p->passthrough ? 1 : 0;
#define RTB_MONO_SPEC(N) fft_kernel_launch<N>(ols_mono_spectrum_kernel<N>, mp, (long long)p->S, st)
    switch (p->N2) {
        case 64: RTB_MONO_SPEC(64); break;   case 128: RTB_MONO_SPEC(128); break;
        case 256: RTB_MONO_SPEC(256); break; default: RTB_MONO_SPEC(512); break;
    }
#undef RTB_MONO_SPEC
    p->launches++;
    if (p->passthrough) {
        ols_mono_passthrough_kernel<<<1184, 256, 0, st>>>(d_in, d_out, p->S, p->Cout, p->B);
        p->launches++;
        return RTB_OK;
    }
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, p->device);
    const long long tiles = (long long)F * Spad / RTB_TC_M;
    const int grid = (int)std::min<long long>(tiles, n_sm);
    kernel_attrs_once(sh_analysis_tc_kernel<true, RTB_TC_LG, true>, 224 * 1024);
    for (int ch = 0, c0 = 0; ch < p->tc_nchunks; c0 += p->tc_nch[ch], ++ch) {
        const int Npad = p->tc_Npad[ch];
        sh_analysis_tc_kernel<true, RTB_TC_LG, true><<<grid, RTB_TC_THREADS, p->tc_smem[ch], st>>>(
            nxt, p->d_bmat[ch], p->d_Y[ch], F, Kpad, p->tc_Jc[ch], Spad, Kpad, Npad, p->tc_cols[ch], 2,
            (long long)2 * Kpad * Npad, 2 * head, 2 * p->P);
        OlsMonoIfftParams ip;
        ip.Y = p->d_Y[ch]; ip.x = d_in; ip.tw = p->d_tw; ip.out = d_out;
        ip.S = p->S; ip.Spad = Spad; ip.Jc = p->tc_Jc[ch]; ip.Cout = p->Cout; ip.c0 = c0; ip.nch = p->tc_nch[ch];
        ip.passthrough = 0;
        const dim3 g2((p->S + RTB_OLS_MONO_ST - 1) / RTB_OLS_MONO_ST, (p->tc_nch[ch] + 1) / 2);
        const size_t smem = sizeof(float) * (size_t)F * 4 * (RTB_OLS_MONO_ST + 1);
#define RTB_MONO_IFFT(N)                                                                             \
    {                                                                                                \
        if (smem > 48 * 1024) kernel_attrs_once(ols_mono_ifft_kernel<N>, (int)smem);                 \
        ols_mono_ifft_kernel<N><<<g2, 128, smem, st>>>(ip);                                          \
    }
        switch (p->N2) {
            case 64: RTB_MONO_IFFT(64) break;   case 128: RTB_MONO_IFFT(128) break;
            case 256: RTB_MONO_IFFT(256) break; default: RTB_MONO_IFFT(512) break;
        }
#undef RTB_MONO_IFFT
        p->launches += 2;
    }
    return RTB_OK;
}
}  // namespace

extern "C" {

int rtb_ols_plan_create(rtb_ols_plan **out, int device, int n_streams, int block_length, int n_in,
                        int n_out, int n_partitions) {
    if (!out) return fail(RTB_ERR_INVALID, "plan output pointer is NULL");
    *out = nullptr;
    if (n_streams < 1 || n_out < 1 || n_partitions < 1)
        return fail(RTB_ERR_INVALID, "n_streams, n_out and n_partitions must be >= 1");
    if (n_in != n_out && n_in != 1)
        return fail(RTB_ERR_INVALID, "channel-wise convolver: n_in must equal n_out or be 1 (got %d, %d)", n_in, n_out);
    if (!block_length_supported(block_length))
        return fail(RTB_ERR_UNSUPPORTED, "block length %d not supported (supported: " RTB_BLOCK_LENGTHS ")", block_length);
    int ndev = 0;
    RTB_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(RTB_ERR_INVALID, "no CUDA device %d", device);
    rtb_ols_plan *p = new (std::nothrow) rtb_ols_plan();
    if (!p) return fail(RTB_ERR_NOMEM, "out of host memory");
    p->device = device; p->S = n_streams; p->B = block_length; p->Cin = n_in; p->Cout = n_out;
    p->P = n_partitions; p->F = block_length + 1; p->Fp = (p->F + 3) / 4 * 4; p->N2 = 2 * block_length;
    DeviceGuard guard(device);
    cudaError_t err = cudaStreamCreateWithFlags(&p->own_stream, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaMalloc(&p->d_xstate, sizeof(float) * 2 * state_elems(p));
    if (err == cudaSuccess) err = cudaMalloc(&p->d_hfd, sizeof(float2) * (size_t)p->P * p->Cout * p->Fp);
    if (err == cudaSuccess) err = cudaMalloc(&p->d_tw, sizeof(float2) * p->N2);
    if (err != cudaSuccess) {
        rtb_ols_plan_destroy(p);
        return fail(err == cudaErrorMemoryAllocation ? RTB_ERR_NOMEM : RTB_ERR_CUDA,
                    "plan allocation failed: %s", cudaGetErrorString(err));
    }
    {   // mono-in / many-out: per-bin GEMM on the tensor cores (RTB_OLS_TC=0 keeps the SIMT kernel)
        const char *env = getenv("RTB_OLS_TC");
        const bool want = env ? env[0] != '0' : true;
        if (want && p->Cin == 1 && p->Cout >= 32 && p->N2 <= 512) {
            int spad = 128;
            while (spad < p->S) spad *= 2;             // tile rows = streams: power of two, >= one M = 128 tile
            p->tc_Spad = spad;
            p->tc_Kpad = (2 * p->P + RTB_TC_KC - 1) / RTB_TC_KC * RTB_TC_KC;
            const int per_chunk = 112;                  // channels per launch: N = 224 <= 256 TMEM columns
            p->tc_nchunks = (p->Cout + per_chunk - 1) / per_chunk;
            if (p->tc_nchunks <= 8) {
                p->tc_mono = true;
                const size_t hbytes = sizeof(float) * (size_t)p->F * p->tc_Kpad * spad;
                err = cudaMalloc(&p->d_histT[0], hbytes);
                for (int ch = 0; ch < p->tc_nchunks && err == cudaSuccess; ++ch) {
                    const int nch = std::min(per_chunk, p->Cout - ch * per_chunk);
                    p->tc_nch[ch] = nch;
                    p->tc_Jc[ch] = 2 * nch;
                    p->tc_Npad[ch] = (2 * nch + 15) / 16 * 16;
                    const int np_ = p->tc_Npad[ch];
                    p->tc_cols[ch] = np_ <= 32 ? 32 : (np_ <= 64 ? 64 : (np_ <= 128 ? 128 : 256));
                    p->tc_smem[ch] = 2 * (int)sizeof(float) * (2 * RTB_TC_KC * RTB_TC_M + 2 * RTB_TC_KC * np_);
                    err = cudaMalloc(&p->d_bmat[ch], sizeof(float) * (size_t)p->F * 2 * p->tc_Kpad * np_);
                    if (err == cudaSuccess)
                        err = cudaMalloc(&p->d_Y[ch], sizeof(float) * (size_t)p->F * p->tc_Jc[ch] * spad);
                }
                if (err != cudaSuccess) {
                    // the spectra buffers of very large batches may not fit: keep the SIMT kernel
                    for (auto &q : p->d_histT) { cudaFree(q); q = nullptr; }
                    for (auto &q : p->d_bmat) { cudaFree(q); q = nullptr; }
                    for (auto &q : p->d_Y) { cudaFree(q); q = nullptr; }
                    p->tc_mono = false;
                    (void)cudaGetLastError();
                    err = cudaSuccess;
                }
            }
        }
    }
    if (!p->tc_mono) {  // delay line and partition mask of the FDL kernel (the tensor-core path has its own line)
        err = cudaMalloc(&p->d_hist, sizeof(float2) * hist_elems(p));
        if (err == cudaSuccess) err = cudaMalloc(&p->d_incl, p->P);
        if (err != cudaSuccess) {
            rtb_ols_plan_destroy(p);
            return fail(err == cudaErrorMemoryAllocation ? RTB_ERR_NOMEM : RTB_ERR_CUDA,
                        "plan allocation failed: %s", cudaGetErrorString(err));
        }
    }
    std::vector<float2> tw(p->N2);
    for (int q = 0; q < p->N2; ++q) {
        const double a = -2.0 * M_PI * q / p->N2;
        tw[q] = make_float2((float)cos(a), (float)sin(a));
    }
    RTB_CUDA(cudaMemcpy(p->d_tw, tw.data(), sizeof(float2) * p->N2, cudaMemcpyHostToDevice));
    *out = p;
    return rtb_ols_plan_reset(p);
}

int rtb_ols_plan_set_filter(rtb_ols_plan *p, const float *hfd) {
    if (!p || !hfd) return fail(RTB_ERR_INVALID, "NULL argument");
    DeviceGuard guard(p->device);
    std::vector<float2> h((size_t)p->P * p->Cout * p->Fp, make_float2(0.f, 0.f));
    for (int q = 0; q < p->P; ++q)
        for (int c = 0; c < p->Cout; ++c)
            for (int f = 0; f < p->F; ++f) {
                const size_t i = ((size_t)q * p->Cout + c) * p->F + f;
                h[((size_t)q * p->Cout + c) * p->Fp + f] = make_float2(hfd[2 * i], hfd[2 * i + 1]);
            }
    RTB_CUDA(cudaDeviceSynchronize());
    RTB_CUDA(cudaMemcpy(p->d_hfd, h.data(), sizeof(float2) * h.size(), cudaMemcpyHostToDevice));
    if (p->tc_mono) {
        // B operand of bin f, K-major core-matrix layout [k/4][n/8][8][4], split into TF32 hi / lo:
        // row n = (channel, re / im), column k = (age p, re / im):
        //   Yr = sum_p Hr Xr - Hi Xi,   Yi = sum_p Hi Xr + Hr Xi
        const int Kpad = p->tc_Kpad;
        for (int ch = 0, c0 = 0; ch < p->tc_nchunks; c0 += p->tc_nch[ch], ++ch) {
            const int Npad = p->tc_Npad[ch];
            const size_t per_f = (size_t)2 * Kpad * Npad;
            std::vector<float> bm((size_t)p->F * per_f, 0.f);
            for (int f = 0; f < p->F; ++f)
                for (int cl = 0; cl < p->tc_nch[ch]; ++cl)
                    for (int q = 0; q < p->P; ++q) {
                        const float2 hv = h[((size_t)q * p->Cout + c0 + cl) * p->Fp + f];
                        const float vals[2][2] = {{hv.x, -hv.y}, {hv.y, hv.x}};  // [n re/im][k re/im]
                        for (int nr = 0; nr < 2; ++nr)
                            for (int kr = 0; kr < 2; ++kr) {
                                const int n = 2 * cl + nr, k = 2 * q + kr;
                                const float v = vals[nr][kr];
                                uint32_t u;
                                memcpy(&u, &v, 4);
                                u = (u + 0x1000u) & 0xffffe000u;  // round to nearest TF32
                                float hi;
                                memcpy(&hi, &u, 4);
                                const size_t off = (((size_t)(k / 4) * (Npad / 8) + n / 8) * 8 + n % 8) * 4 + k % 4;
                                bm[f * per_f + off] = hi;
                                bm[f * per_f + (size_t)Kpad * Npad + off] = v - hi;
                            }
                    }
            RTB_CUDA(cudaMemcpy(p->d_bmat[ch], bm.data(), sizeof(float) * bm.size(), cudaMemcpyHostToDevice));
        }
    }
    RTB_CUDA(cudaDeviceSynchronize());  // uploads and creation-time memsets done before a non-blocking stream runs
    p->filter_set = true;
    return RTB_OK;
}

int rtb_ols_plan_set_passthrough(rtb_ols_plan *p, int on) {
    if (!p) return fail(RTB_ERR_INVALID, "NULL plan");
    p->passthrough = on != 0;
    return RTB_OK;
}

int rtb_ols_plan_reset(rtb_ols_plan *p) {
    if (!p) return fail(RTB_ERR_INVALID, "NULL plan");
    DeviceGuard guard(p->device);
    // _input_block_td.fill(0); _blocks_fd.fill(0)  (_convolver.py:500-505).  The memsets run on the legacy
    // default stream, which does not order against the plan's (non-blocking) streams or a caller's:
    // wait for the blocks in flight first and for the memsets afterwards.
    RTB_CUDA(cudaDeviceSynchronize());
    RTB_CUDA(cudaMemset(p->d_xstate, 0, sizeof(float) * 2 * state_elems(p)));
    if (p->d_hist) RTB_CUDA(cudaMemset(p->d_hist, 0, sizeof(float2) * hist_elems(p)));
    if (p->d_incl) RTB_CUDA(cudaMemset(p->d_incl, 0, p->P));
    if (p->tc_mono) {
        RTB_CUDA(cudaMemset(p->d_histT[0], 0, sizeof(float) * (size_t)p->F * p->tc_Kpad * p->tc_Spad));
        p->tc_hpar = 0;
    }
    RTB_CUDA(cudaDeviceSynchronize());
    return RTB_OK;
}

int rtb_ols_process_block(rtb_ols_plan *p, const float *d_in, float *d_out, void *cuda_stream) {
    if (!p || !d_in || !d_out) return fail(RTB_ERR_INVALID, "NULL argument");
    if (!p->filter_set) return fail(RTB_ERR_STATE, "blocks in frequency domain have not been calculated yet.");
    DeviceGuard guard(p->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    OlsParams prm;
    prm.x = d_in;
    prm.xprev = p->d_xstate + (size_t)p->par * state_elems(p);
    prm.xnext = p->d_xstate + (size_t)(p->par ^ 1) * state_elems(p);
    prm.hist = p->d_hist; prm.hfd = p->d_hfd; prm.incl = p->d_incl; prm.tw = p->d_tw; prm.out = d_out;
    prm.S = p->S; prm.Cin = p->Cin; prm.Cout = p->Cout; prm.P = p->P; prm.Fp = p->Fp;
    prm.slot = p->slot; prm.passthrough = p->passthrough ? 1 : 0;
    const int npairs = (p->Cout + 1) / 2;
    if (p->profiling) cudaEventRecord(p->ev[0], st);
    if (p->tc_mono) {
        const int rc = process_mono_tc(p, d_in, d_out, st);
        if (rc != RTB_OK) return rc;
        if (p->profiling) { cudaEventRecord(p->ev[1], st); p->ev_valid = true; }
        p->par ^= 1;
        RTB_CUDA(cudaGetLastError());
        return RTB_OK;
    }
    // tile shape: mono-in / many-out (ARIR) shares both operands -> square tile; otherwise all
    // transforms of a CTA take the same pair of consecutive streams
#define RTB_OLS_LAUNCH(N)                                                                            \
    {                                                                                                \
        const int tpc = FftCfg<N>::TASKS;                                                            \
        const bool mono = p->Cin == 1 && p->Cout > 1;                                                \
        prm.TP = mono ? std::min(npairs, tpc >= 16 ? 4 : (tpc >= 4 ? 2 : 1)) : 1;                    \
        prm.TS = tpc / prm.TP;                                                                       \
        prm.tiles_s = (p->S + prm.TS - 1) / prm.TS;                                                  \
        const long long tiles = (long long)prm.tiles_s * ((npairs + prm.TP - 1) / prm.TP);          \
        fft_kernel_launch<N>(ols_fdl_kernel<N>, prm, tiles * tpc, st);                               \
    }
    RTB_DISPATCH_N2(p->N2, RTB_OLS_LAUNCH)
#undef RTB_OLS_LAUNCH
    if (p->profiling) { cudaEventRecord(p->ev[1], st); p->ev_valid = true; }
    p->launches++;
    p->slot = (p->slot + 1) % p->P;
    p->par ^= 1;
    RTB_CUDA(cudaGetLastError());
    return RTB_OK;
}

int rtb_ols_process_block_host(rtb_ols_plan *p, const float *h_in, float *h_out) {
    if (!p || !h_in || !h_out) return fail(RTB_ERR_INVALID, "NULL argument");
    DeviceGuard guard(p->device);
    const size_t n_in = state_elems(p), n_out = (size_t)p->S * p->Cout * p->B;
    if (!p->d_in) {
        cudaError_t err = cudaMalloc(&p->d_in, sizeof(float) * n_in);
        if (err == cudaSuccess) err = cudaMalloc(&p->d_out, sizeof(float) * n_out);
        if (err != cudaSuccess) return fail(RTB_ERR_NOMEM, "staging allocation failed: %s", cudaGetErrorString(err));
    }
    RTB_CUDA(cudaMemcpyAsync(p->d_in, h_in, sizeof(float) * n_in, cudaMemcpyHostToDevice, p->own_stream));
    int rc = rtb_ols_process_block(p, p->d_in, p->d_out, p->own_stream);
    if (rc != RTB_OK) return rc;
    RTB_CUDA(cudaMemcpyAsync(h_out, p->d_out, sizeof(float) * n_out, cudaMemcpyDeviceToHost, p->own_stream));
    RTB_CUDA(cudaStreamSynchronize(p->own_stream));
    return RTB_OK;
}

int rtb_ols_plan_set_profiling(rtb_ols_plan *p, int on) {
    if (!p) return fail(RTB_ERR_INVALID, "NULL plan");
    DeviceGuard guard(p->device);
    if (on && !p->ev[0])
        for (auto &e : p->ev) RTB_CUDA(cudaEventCreate(&e));
    p->profiling = on != 0;
    p->ev_valid = false;
    return RTB_OK;
}

int rtb_ols_plan_query(rtb_ols_plan *p, int what, int64_t *value) {
    if (!p || !value) return fail(RTB_ERR_INVALID, "NULL argument");
    switch (what) {
        case RTB_Q_LAUNCHES: *value = p->launches; break;
        case RTB_Q_STATE_BYTES:
            *value = (int64_t)sizeof(float2) * p->P * p->Cin * p->Fp + (int64_t)sizeof(float) * 2 * p->Cin * p->B;
            break;
        case RTB_Q_RENDER_NS: {
            if (!p->profiling || !p->ev_valid) return fail(RTB_ERR_STATE, "profiling is off or no block was processed");
            DeviceGuard guard(p->device);
            RTB_CUDA(cudaEventSynchronize(p->ev[1]));
            float ms = 0.f;
            RTB_CUDA(cudaEventElapsedTime(&ms, p->ev[0], p->ev[1]));
            *value = (int64_t)(ms * 1e6);
            break;
        }
        default: return fail(RTB_ERR_INVALID, "unknown query %d", what);
    }
    return RTB_OK;
}

int rtb_ols_plan_destroy(rtb_ols_plan *p) {
    if (!p) return RTB_OK;
    DeviceGuard guard(p->device);
    cudaFree(p->d_xstate); cudaFree(p->d_hist); cudaFree(p->d_hfd); cudaFree(p->d_incl); cudaFree(p->d_tw);
    cudaFree(p->d_in); cudaFree(p->d_out);
    for (auto q : p->d_histT) cudaFree(q);
    for (auto q : p->d_bmat) cudaFree(q);
    for (auto q : p->d_Y) cudaFree(q);
    if (p->own_stream) cudaStreamDestroy(p->own_stream);
    for (auto &e : p->ev) if (e) cudaEventDestroy(e);
    delete p;
    return RTB_OK;
}

}  // extern "C"
