// Device kernels of the SH-domain renderer (reference: AdjustableShConvolver.filter_block,
// ReTiSAR/_convolver.py:1231-1316).  See DESIGN.md for the derivation; in short
//
//   K-A  sh_analysis_kernel : z[j][t] = sum_c Rm[j][c] * x[c][t]           (time-domain SFT)
//   K-B  sh_render_kernel   : FFT of the complex signals built from z rows (overlap-save over the
//        previous and current block), SH-domain HRIR multiply + head rotation + sum over SH
//        coefficients for the current and the previous yaw, two inverse FFTs, overlap-save
//        discard and cos^2 cross-fade -- all in registers of one warp per stream.
//
// Because the spherical Fourier transform matrix Y_w is frequency independent, it commutes with
// the FFT (_convolver.py:1260 applies it after the rFFT, here it is applied before); with the
// (n,-m) conjugate symmetry of Y_w only K real rows are needed and the transform of the complex
// signal s_(n,m)(t) yields both S_(n,m)(f) (bin f) and S_(n,-m)(f) = (-1)^m conj(bin N2-f).
#pragma once
#include <cuda_fp16.h>

#include "fft_warp.cuh"

namespace rtb {

// ---------------------------------------------------------------------------------------------
// K-A: batched real GEMM, one warp per (stream, 32*TT samples, JT rows) tile, no shared memory:
// x is read once per tile with coalesced vector loads, the analysis matrix through broadcast
// loads that stay L1-resident.
// ---------------------------------------------------------------------------------------------
template <int TT> struct VecT;
template <> struct VecT<1> { using type = float; };
template <> struct VecT<2> { using type = float2; };
template <> struct VecT<4> { using type = float4; };

template <int JT, int TT>
__global__ void __launch_bounds__(128)
sh_analysis_kernel(const float *__restrict__ x,    // [S][C][B]
                   const float *__restrict__ rmT,  // [C][Jpad] analysis matrix, transposed
                   float *__restrict__ zcur,       // [S][J][B]
                   int S, int C, int J, int Jpad, int B, int n_ttiles, int n_jchunks) {
    using V = typename VecT<TT>::type;
    const int lane = threadIdx.x & 31;
    const long long task = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    const long long total = (long long)S * n_ttiles * n_jchunks;
    if (task >= total) return;
    const int jc = (int)(task % n_jchunks);
    const long long rest = task / n_jchunks;
    const int tt = (int)(rest % n_ttiles);
    const int s = (int)(rest / n_ttiles);
    const int t0 = tt * 32 * TT + lane * TT;
    const int j0 = jc * JT;

    const float *xp = x + (size_t)s * C * B + t0;
    const float *rp = rmT + j0;
    float acc[JT][TT];
#pragma unroll
    for (int j = 0; j < JT; ++j)
#pragma unroll
        for (int i = 0; i < TT; ++i) acc[j][i] = 0.f;

#pragma unroll 4
    for (int c = 0; c < C; ++c) {
        V xv = __ldg(reinterpret_cast<const V *>(xp + (size_t)c * B));
        const float *xs = reinterpret_cast<const float *>(&xv);
        float r[JT];
#pragma unroll
        for (int j4 = 0; j4 < JT / 4; ++j4) {
            float4 rv = __ldg(reinterpret_cast<const float4 *>(rp + (size_t)c * Jpad) + j4);
            r[4 * j4] = rv.x; r[4 * j4 + 1] = rv.y; r[4 * j4 + 2] = rv.z; r[4 * j4 + 3] = rv.w;
        }
#pragma unroll
        for (int j = 0; j < JT; ++j)
#pragma unroll
            for (int i = 0; i < TT; ++i) acc[j][i] = fmaf(r[j], xs[i], acc[j][i]);
    }
#pragma unroll
    for (int j = 0; j < JT; ++j) {
        if (j0 + j < J) {
            V v;
            float *vs = reinterpret_cast<float *>(&v);
#pragma unroll
            for (int i = 0; i < TT; ++i) vs[i] = acc[j][i];
            *reinterpret_cast<V *>(zcur + ((size_t)s * J + j0 + j) * B + t0) = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// yaw (degrees) -> float16-quantised radians, bit-for-bit the reference's chain
// (_convolver.py:729, :947-975, :1279-1281; SURVEY.md section 8a "exact yaw -> phase recipe")
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float rn16(float v) { return __half2float(__float2half_rn(v)); }

__device__ __forceinline__ float yaw_deg_to_rad16(float yaw_deg, float tracker_dir, float src_azim16) {
    float a = rn16(tracker_dir * yaw_deg);  // python float is a weak scalar -> float16
    a = rn16(a + src_azim16);               // + _sources_deg[:, 0] (float16)
    a = rn16(a + 360.0f);
    float r = fmodf(a, 360.0f);             // numpy remainder: python-style, result in [0, 360)
    if (r != 0.0f) {
        if (r < 0.0f) r += 360.0f;
    } else {
        r = 0.0f;
    }
    r = rn16(r);                            // exact, kept for symmetry with the half loop
    return rn16(r * 0.017453292519943295f); // np.deg2rad on float16
}

// ---------------------------------------------------------------------------------------------
// K-B
// ---------------------------------------------------------------------------------------------
struct ShSignal {   // one complex FFT signal (see sh_plan.cu: build_signals)
    int row_re;     // analysis row giving the real part
    int row_im;     // analysis row giving the imaginary part, -1 = zero
    int m1;         // term 1: bin f,              phase exp(-i*m1*alpha)
    int m2;         // term 2: conj(bin N2-f),     phase exp(-i*m2*alpha); RTB_NO_TERM = absent
};
#define RTB_NO_TERM 0x7fffffff
#define RTB_MAX_ORDER 31

struct ShRenderParams {
    const float *zprev;          // [S][J][B] analysis rows of the previous block
    const float *zcur;           // [S][J][B] analysis rows of the current block
    const float4 *htab;          // [NP][2][F] (ear0.re, ear0.im, ear1.re, ear1.im)
    const ShSignal *signals;     // [NP]
    const float2 *tw;            // [N2] exp(-2 pi i q / N2)
    const float *win_in;         // [B] fade-in window  / N2
    const float *win_out;        // [B] fade-out window / N2
    const float *yaw_deg;        // [S]
    const float *rad_last_in;    // [S] float16-quantised radians of the previous block
    float *rad_last_out;         // [S] (written when cross-fading)
    float *out;                  // [S][2][B]
    int S, J, B;
    int np_cur;                  // signals (prefix) taking part in the current-yaw sum
    int np_last;                 // ... in the previous-yaw sum (0 = none: first block / fade off)
    int order_max;               // SH order (for the phase tables)
    int crossfade;
    float tracker_dir, src_azim16;
};

// phase = exp(-i * m * alpha) from the (cos, sin)(|m| alpha) table
__device__ __forceinline__ float2 phase_of(const float2 *cs, int m) {
    float2 v = cs[m < 0 ? -m : m];
    return make_float2(v.x, m < 0 ? v.y : -v.y);
}

template <int N2>
__global__ void __launch_bounds__(128)
sh_render_kernel(const ShRenderParams prm) {
    using Cfg = FftCfg<N2>;
    constexpr int G = Cfg::G, EPL = Cfg::EPL, HB = EPL / 2, GROUPS = Cfg::GROUPS;
    constexpr int B = N2 / 2, F = B + 1;
    constexpr unsigned FULL = 0xffffffffu;

    __shared__ float2 s_tw[N2];
    __shared__ float2 s_buf[4][GROUPS * Cfg::REGION];
    __shared__ float2 s_cs[4][2][RTB_MAX_ORDER + 1];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane / G, l = lane % G;
    for (int i = threadIdx.x; i < N2; i += 128) s_tw[i] = prm.tw[i];
    __syncthreads();

    const int s = blockIdx.x * 4 + warp;
    if (s >= prm.S) return;
    float2 *buf = s_buf[warp] + grp * Cfg::REGION;
    const int partner = (G - l) & (G - 1);

    // rotation phases for the current and the previous yaw (a4)
    const float rad_cur = yaw_deg_to_rad16(prm.yaw_deg[s], prm.tracker_dir, prm.src_azim16);
    const float rad_last = prm.rad_last_in[s];
    if (lane <= prm.order_max) {
        float sn, cs;
        sincosf((float)lane * rad_cur, &sn, &cs);
        s_cs[warp][0][lane] = make_float2(cs, sn);
        sincosf((float)lane * rad_last, &sn, &cs);
        s_cs[warp][1][lane] = make_float2(cs, sn);
    }
    __syncwarp();
    if (prm.crossfade && lane == 0) prm.rad_last_out[s] = rad_cur;

    // accumulators: [bin slot][cur ear0, cur ear1, last ear0, last ear1]
    float2 acc[HB + 1][4];
#pragma unroll
    for (int u = 0; u <= HB; ++u)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[u][q] = make_float2(0.f, 0.f);

    const size_t zoff = (size_t)s * prm.J * B;
    const int np = max(prm.np_cur, prm.np_last);
    for (int p0 = 0; p0 < np; p0 += GROUPS) {
        const int p = p0 + grp;
        const bool valid = p < np;
        ShSignal sg = prm.signals[valid ? p : 0];
        // ---- a1: overlap-save buffer (previous | current block) of the complex signal
        float2 X[EPL];
        {
            const float *re_prev = prm.zprev + zoff + (size_t)sg.row_re * B + l;
            const float *re_cur = prm.zcur + zoff + (size_t)sg.row_re * B + l;
#pragma unroll
            for (int u = 0; u < HB; ++u) {
                X[u].x = valid ? __ldg(re_prev + G * u) : 0.f;
                X[u + HB].x = valid ? __ldg(re_cur + G * u) : 0.f;
            }
            if (valid && sg.row_im >= 0) {
                const float *im_prev = prm.zprev + zoff + (size_t)sg.row_im * B + l;
                const float *im_cur = prm.zcur + zoff + (size_t)sg.row_im * B + l;
#pragma unroll
                for (int u = 0; u < HB; ++u) {
                    X[u].y = __ldg(im_prev + G * u);
                    X[u + HB].y = __ldg(im_cur + G * u);
                }
            } else {
#pragma unroll
                for (int u = 0; u < EPL; ++u) X[u].y = 0.f;
            }
        }
        warp_fft<N2, false>(X, buf, s_tw, l);

        // ---- a3-a5: HRIR multiply, rotation, sum over SH coefficients
        const bool in_cur = valid && p < prm.np_cur, in_last = valid && p < prm.np_last;
        float2 ph1c = make_float2(0.f, 0.f), ph1l = ph1c, ph2c = ph1c, ph2l = ph1c;
        const bool has2 = sg.m2 != RTB_NO_TERM;
        if (in_cur) {
            ph1c = phase_of(s_cs[warp][0], sg.m1);
            if (has2) ph2c = phase_of(s_cs[warp][0], sg.m2);
        }
        if (in_last) {
            ph1l = phase_of(s_cs[warp][1], sg.m1);
            if (has2) ph2l = phase_of(s_cs[warp][1], sg.m2);
        }
        const float4 *h1 = prm.htab + (size_t)(valid ? p : 0) * 2 * F;
        const float4 *h2 = h1 + F;
#pragma unroll
        for (int u = 0; u <= HB; ++u) {
            const int f = (u < HB) ? (l + G * u) : B;  // slot HB is the Nyquist bin (lane 0 only)
            const float2 zf = X[u];
            {
                const float4 h = __ldg(h1 + f);
                const float2 tc = cmul(ph1c, zf), tl = cmul(ph1l, zf);
                cfma(acc[u][0], make_float2(h.x, h.y), tc);
                cfma(acc[u][1], make_float2(h.z, h.w), tc);
                cfma(acc[u][2], make_float2(h.x, h.y), tl);
                cfma(acc[u][3], make_float2(h.z, h.w), tl);
            }
            // element N2-f: slot EPL-1-u of the partner lane (lane 0: own slot EPL-u, bin 0: itself)
            float2 zp;
            if (u < HB) {
                zp.x = __shfl_sync(FULL, X[EPL - 1 - u].x, partner, G);
                zp.y = __shfl_sync(FULL, X[EPL - 1 - u].y, partner, G);
                if (l == 0) zp = (u == 0) ? X[0] : X[(EPL - u) % EPL];
            } else {
                zp = X[HB];
            }
            if (has2) {  // uniform inside a group; groups diverge only on the shuffle-free part
                const float4 h = __ldg(h2 + f);
                const float2 zc = make_float2(zp.x, -zp.y);
                const float2 tc = cmul(ph2c, zc), tl = cmul(ph2l, zc);
                cfma(acc[u][0], make_float2(h.x, h.y), tc);
                cfma(acc[u][1], make_float2(h.z, h.w), tc);
                cfma(acc[u][2], make_float2(h.x, h.y), tl);
                cfma(acc[u][3], make_float2(h.z, h.w), tl);
            }
        }
    }

    // transforms that ran side by side in one warp each hold a partial sum
#pragma unroll
    for (int off = G; off < 32; off <<= 1) {
#pragma unroll
        for (int u = 0; u <= HB; ++u)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                acc[u][q].x += __shfl_xor_sync(FULL, acc[u][q].x, off);
                acc[u][q].y += __shfl_xor_sync(FULL, acc[u][q].y, off);
            }
    }

    // ---- a6, a7: Hermitian extension, inverse FFT of (left + i right), discard, cross-fade.
    // With G < 32 the current-yaw pair is transformed by group 0 and the previous-yaw pair by
    // group 1 at the same time; with G == 32 the warp does them one after the other.
    float2 ycur[HB], ylast[HB];
    constexpr int NITER = (GROUPS >= 2) ? 1 : 2;
#pragma unroll
    for (int it = 0; it < NITER; ++it) {
        const bool use_last = (GROUPS >= 2) ? (grp == 1) : (it == 1);
        float2 X[EPL], V[HB];
#pragma unroll
        for (int u = 0; u < HB; ++u) {
            float2 eL = use_last ? acc[u][2] : acc[u][0];
            float2 eR = use_last ? acc[u][3] : acc[u][1];
            if (u == 0 && l == 0) { eL.y = 0.f; eR.y = 0.f; }  // irfft ignores Im of the DC bin
            X[u] = make_float2(eL.x - eR.y, eL.y + eR.x);      // Z(f)    = L + i R
            V[u] = make_float2(eL.x + eR.y, eR.x - eL.y);      // Z(N2-f) = conj(L) + i conj(R)
        }
        const float2 nyq = make_float2(use_last ? acc[HB][2].x : acc[HB][0].x,
                                       use_last ? acc[HB][3].x : acc[HB][1].x);
#pragma unroll
        for (int u2 = HB; u2 < EPL; ++u2) {
            float2 x;
            x.x = __shfl_sync(FULL, V[EPL - 1 - u2].x, partner, G);
            x.y = __shfl_sync(FULL, V[EPL - 1 - u2].y, partner, G);
            if (l == 0) x = (u2 == HB) ? nyq : V[(EPL - u2) % HB];
            X[u2] = x;
        }
        warp_fft<N2, true>(X, buf, s_tw, l);
#pragma unroll
        for (int u = 0; u < HB; ++u) {
            if (GROUPS >= 2 || it == 0) ycur[u] = X[u + HB];
            if (GROUPS >= 2 || it == 1) ylast[u] = X[u + HB];
        }
    }
    if (GROUPS >= 2) {  // bring the previous-yaw block from group 1 to group 0
#pragma unroll
        for (int u = 0; u < HB; ++u) {
            ylast[u].x = __shfl_sync(FULL, ylast[u].x, (lane + G) & 31);
            ylast[u].y = __shfl_sync(FULL, ylast[u].y, (lane + G) & 31);
        }
    }
    if (grp == 0) {
        float *o = prm.out + (size_t)s * 2 * B;
#pragma unroll
        for (int u = 0; u < HB; ++u) {
            const int n = l + G * u;
            float yl, yr;
            if (prm.crossfade) {
                const float wi = __ldg(prm.win_in + n), wo = __ldg(prm.win_out + n);
                yl = ycur[u].x * wi + ylast[u].x * wo;
                yr = ycur[u].y * wi + ylast[u].y * wo;
            } else {
                const float sc = 1.0f / (float)N2;
                yl = ycur[u].x * sc;
                yr = ycur[u].y * sc;
            }
            o[n] = yl;
            o[B + n] = yr;
        }
    }
}

// a10: passthrough.  irfft(rfft(buffer))[B:] is the current block itself, so the first E input
// channels are copied to the ears (_convolver.py:558-562); missing channels stay silent.
__global__ void sh_passthrough_kernel(const float *__restrict__ x, float *__restrict__ out,
                                      int S, int C, int E, int B) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)S * E * B;
    if (i >= total) return;
    const int t = (int)(i % B);
    const int e = (int)((i / B) % E);
    const long long s = i / ((long long)B * E);
    out[i] = (e < C) ? x[((size_t)s * C + e) * B + t] : 0.f;
}

}  // namespace rtb
