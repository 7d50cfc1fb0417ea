// Headphone-compensation filter (HPCF) behind the SH renderer, inside the renderer's own launch.
//
// The reference chains the HPCF as a separate JACK client after the renderer
// (`_run.py:290-338`): a 2-in / 2-out channel-wise `OverlapSaveConvolver` (`_convolver.py:524-571`,
// `get_input_channel_count` :409-410) fed with the binaural block.  The cross-fade of the renderer
// happens in the time domain, so the filter cannot move in front of the renderer's inverse FFT; but
// the warp that has just produced a stream's two ear signals can carry on with them: the binaural
// block goes to the HPCF's overlap-save state (which must be stored anyway), the warp transforms
// (previous | current) block of both ears in one complex FFT, appends the spectrum to an input-side
// delay line and accumulates  Y = sum_p H[p] X_(t-p)  exactly like `ols_fdl_kernel` (ols_plan.cu).
// Against a separate plan this saves one launch, the HBM read of the binaural block and the second
// write of the state.  Included by sh_kernels.cuh before the render kernels.
#pragma once

namespace rtb {

struct HpcfParams {
    const float *x_prev;   // [S][2][B] binaural block of the previous call (state, read side)
    float *x_cur;          // [S][2][B] binaural block of this call: written by the renderer, read here
    float2 *hist;          // [S][P][2][Fp] delay line of binaural spectra, ring over P
    const float2 *hfd;     // [P][2][Fp] filter partition spectra
    float *out;            // [S][2][B] what the caller receives
    int P;                 // partitions, 0 = no HPCF attached
    int Fp, slot;
};

// All lanes of the transform's group call it together; only `valid` groups touch memory.
template <int N2>
__device__ __forceinline__ void hpcf_stage(const HpcfParams &hp, int s, bool valid, float2 *buf,
                                           const float2 *tw, int l) {
    using Cfg = FftCfg<N2>;
    constexpr int G = Cfg::G, EPL = Cfg::EPL, HB = EPL / 2, B = N2 / 2;
    // the ear signals were written by other lanes of this transform group a moment ago
    fft_sync<N2>();
    float2 X[EPL];
    const size_t r0 = (size_t)s * 2 * B + l, r1 = r0 + B;
#pragma unroll
    for (int u = 0; u < HB; ++u) {
        X[u].x = valid ? hp.x_prev[r0 + G * u] : 0.f;
        X[u].y = valid ? hp.x_prev[r1 + G * u] : 0.f;
        X[u + HB].x = valid ? __ldcg(hp.x_cur + r0 + G * u) : 0.f;
        X[u + HB].y = valid ? __ldcg(hp.x_cur + r1 + G * u) : 0.f;
    }
    warp_fft<N2, false>(X, buf, tw, l);
    float2 XL[HB + 1], XR[HB + 1], M[HB + 1];
    fft_mirror<N2>(X, M, buf, l);
#pragma unroll
    for (int u = 0; u <= HB; ++u) {  // Z = L + iR with L, R Hermitian
        const float2 zf = X[u], zp = M[u];
        XL[u] = make_float2(0.5f * (zf.x + zp.x), 0.5f * (zf.y - zp.y));
        XR[u] = make_float2(0.5f * (zf.y + zp.y), 0.5f * (zp.x - zf.x));
    }
    const int Fp = hp.Fp;
    const bool own_nyq = (l == 0);
    float2 *h0 = hp.hist + (((size_t)s * hp.P + hp.slot) * 2) * Fp;
    float2 YL[HB + 1], YR[HB + 1];
#pragma unroll
    for (int u = 0; u <= HB; ++u) {
        const int f = l + G * u;
        const bool ok = valid && (u < HB || own_nyq);
        if (ok) {
            h0[f] = XL[u];
            h0[Fp + f] = XR[u];
        }
        const float2 a = ok ? __ldg(hp.hfd + f) : make_float2(0.f, 0.f);
        const float2 b = ok ? __ldg(hp.hfd + Fp + f) : make_float2(0.f, 0.f);
        YL[u] = cmul(a, XL[u]);
        YR[u] = cmul(b, XR[u]);
    }
    int hs = hp.slot;
    for (int p = 1; p < hp.P; ++p) {
        hs = (hs == 0) ? hp.P - 1 : hs - 1;
        const float2 *x0 = hp.hist + (((size_t)s * hp.P + hs) * 2) * Fp;
        const float2 *g0 = hp.hfd + (size_t)p * 2 * Fp;
        float2 xa[HB + 1], xb[HB + 1], ga[HB + 1], gb[HB + 1];
#pragma unroll
        for (int u = 0; u <= HB; ++u) {
            const int f = l + G * u;
            const bool ok = valid && (u < HB || own_nyq);
            xa[u] = ok ? __ldcs(x0 + f) : make_float2(0.f, 0.f);
            xb[u] = ok ? __ldcs(x0 + Fp + f) : make_float2(0.f, 0.f);
            ga[u] = ok ? __ldg(g0 + f) : make_float2(0.f, 0.f);
            gb[u] = ok ? __ldg(g0 + Fp + f) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u <= HB; ++u) {
            cfma(YL[u], ga[u], xa[u]);
            cfma(YR[u], gb[u], xb[u]);
        }
    }
    {
        float2 V[HB];
#pragma unroll
        for (int u = 0; u < HB; ++u) {
            float2 eL = YL[u], eR = YR[u];
            if (u == 0 && l == 0) { eL.y = 0.f; eR.y = 0.f; }
            X[u] = make_float2(eL.x - eR.y, eL.y + eR.x);
            V[u] = make_float2(eL.x + eR.y, eR.x - eL.y);
        }
        fft_hermitian_fill<N2>(X, V, make_float2(YL[HB].x, YR[HB].x), buf, l);
    }
    fft_sync<N2>();
    warp_fft<N2, true>(X, buf, tw, l);
    if (valid) {
        const float sc = 1.0f / (float)N2;
        float *o = hp.out + (size_t)s * 2 * B + l;
#pragma unroll
        for (int u = 0; u < HB; ++u) {
            o[G * u] = X[u + HB].x * sc;
            o[B + G * u] = X[u + HB].y * sc;
        }
    }
}

// Stand-alone form for the paths whose last kernel is not a render kernel (passthrough of the
// renderer, partitioned HRIR): one transform group per stream.
template <int N2>
__global__ void __launch_bounds__(FftCfg<N2>::THREADS)
hpcf_kernel(const HpcfParams hp, const float2 *__restrict__ tw_g, int S) {
    using Cfg = FftCfg<N2>;
    constexpr int G = Cfg::G, GROUPS = Cfg::GROUPS;
    constexpr bool BIG = Cfg::BIG;
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    __shared__ float2 s_tw[BIG ? 1 : N2];
    __shared__ __align__(128) float2 s_buf[BIG ? 1 : 4 * GROUPS * Cfg::REGION];
    const int warp = BIG ? 0 : (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int grp = BIG ? 0 : lane / G, l = BIG ? (int)threadIdx.x : lane % G;
    const float2 *tw = tw_g;
    float2 *buf = reinterpret_cast<float2 *>(dyn_smem);
    if constexpr (!BIG) {
        for (int i = threadIdx.x; i < N2; i += 128) s_tw[i] = tw_g[i];
        __syncthreads();
        tw = s_tw;
        buf = s_buf + (warp * GROUPS + grp) * Cfg::REGION;
    }
    const int s = BIG ? (int)blockIdx.x : ((int)blockIdx.x * 4 + warp) * GROUPS + grp;
    if constexpr (!BIG) {
        if (!__any_sync(0xffffffffu, s < S)) return;
    }
    hpcf_stage<N2>(hp, s < S ? s : 0, s < S, buf, tw, l);
}

}  // namespace rtb
