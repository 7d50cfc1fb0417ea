// Warp-level complex FFT (Stockham autosort, radix 8/4/2) held in registers, with the inter-pass
// exchange through a warp-private, XOR-swizzled shared-memory buffer.
//
// A transform of N2 points is owned by a "group" of G lanes (G = 32 for N2 >= 256, N2/8 below),
// so small transforms run 2 or 4 per warp side by side.  Lane l of a group holds EPL = N2/G
// complex values X[u] <-> element index l + G*u, before the first and after the last pass.  That
// layout is what the callers build directly from coalesced global loads and consume directly in
// registers (no shared-memory round trip for the spectrum): elements u < EPL/2 of the result are
// the bins 0 <= f < N2/2 owned by the lane, and element N2-f lives in lane (G-l)%G.
//
// Convention matches numpy: forward unnormalised exp(-2*pi*i*n*k/N2), inverse exp(+...) also
// unnormalised (the caller folds 1/N2 into its window).
#pragma once
#include "rtb_common.cuh"

namespace rtb {

template <int N2>
struct FftCfg {
    static_assert(N2 == 64 || N2 == 128 || N2 == 256 || N2 == 512, "unsupported FFT length");
    static constexpr int G = (N2 >= 256) ? 32 : N2 / 8;  // lanes per transform
    static constexpr int EPL = N2 / G;                   // complex elements per lane (8 or 16)
    static constexpr int GROUPS = 32 / G;                // transforms per warp
    static constexpr int R_LAST = (N2 == 64) ? 0 : (N2 == 128 ? 2 : (N2 == 256 ? 4 : 8));
    static constexpr int REGION = N2 + (N2 == 64 ? 8 : 0);  // float2 slots per transform buffer
};

// Conflict-free placement for every access pattern of the passes below; found by brute force,
// see tools/swizzle_search.py (2 wavefronts per 64-bit warp access = the minimum).
template <int N2>
__host__ __device__ __forceinline__ constexpr int fft_swz(int i) {
    return (N2 == 64) ? (i ^ ((i >> 3) & 7)) : (i ^ ((i >> 3) & 1) ^ (((i >> 4) & 7) << 1));
}
// The swizzle is linear over GF(2), and every index used below is a sum of a lane-dependent part
// and a compile-time part that occupy disjoint bit fields, so
//   fft_swz(lane_part + const_part) == fft_swz(lane_part) ^ fft_swz(const_part)
// : one XOR with an immediate per access instead of recomputing the swizzle.

// multiply by -i (forward) / +i (inverse)
template <bool INV>
__device__ __forceinline__ float2 rot90(float2 a) {
    return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

template <bool INV>
__device__ __forceinline__ void dft2(float2 &b0, float2 &b1) {
    float2 t = b0;
    b0 = cadd(t, b1);
    b1 = csub(t, b1);
}

template <bool INV>
__device__ __forceinline__ void dft4(float2 &b0, float2 &b1, float2 &b2, float2 &b3) {
    float2 c0 = cadd(b0, b2), c1 = csub(b0, b2), c2 = cadd(b1, b3), c3 = rot90<INV>(csub(b1, b3));
    b0 = cadd(c0, c2);
    b1 = cadd(c1, c3);
    b2 = csub(c0, c2);
    b3 = csub(c1, c3);
}

template <bool INV>
__device__ __forceinline__ void dft8(float2 (&v)[8]) {
    const float h = 0.70710678118654752440f;
    float2 a0 = cadd(v[0], v[4]), a1 = cadd(v[1], v[5]), a2 = cadd(v[2], v[6]), a3 = cadd(v[3], v[7]);
    float2 d0 = csub(v[0], v[4]), e1 = csub(v[1], v[5]), e2 = csub(v[2], v[6]), e3 = csub(v[3], v[7]);
    float2 d1, d3;
    if (INV) {
        d1 = make_float2(h * (e1.x - e1.y), h * (e1.x + e1.y));
        d3 = make_float2(h * (-e3.x - e3.y), h * (e3.x - e3.y));
    } else {
        d1 = make_float2(h * (e1.x + e1.y), h * (e1.y - e1.x));
        d3 = make_float2(h * (e3.y - e3.x), h * (-e3.x - e3.y));
    }
    float2 d2 = rot90<INV>(e2);
    dft4<INV>(a0, a1, a2, a3);
    dft4<INV>(d0, d1, d2, d3);
    v[0] = a0; v[2] = a1; v[4] = a2; v[6] = a3;
    v[1] = d0; v[3] = d1; v[5] = d2; v[7] = d3;
}

template <int R, bool INV>
__device__ __forceinline__ void dft_r(float2 (&v)[R]) {
    if constexpr (R == 8) dft8<INV>(v);
    else if constexpr (R == 4) dft4<INV>(v[0], v[1], v[2], v[3]);
    else dft2<INV>(v[0], v[1]);
}

// One Stockham pass of radix R over sub-transforms of length NS (product of earlier radices).
// TO_SMEM: scatter the results into the exchange buffer, otherwise keep them in X (last pass).
template <int N2, int R, int NS, bool INV, bool TO_SMEM>
__device__ __forceinline__ void fft_pass(float2 (&X)[FftCfg<N2>::EPL], float2 *buf,
                                         const float2 *tw, int l) {
    using Cfg = FftCfg<N2>;
    constexpr int NB = Cfg::EPL / R;  // butterflies per lane
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        float2 v[R];
#pragma unroll
        for (int t = 0; t < R; ++t) v[t] = X[b + t * NB];
        const int j = l + Cfg::G * b;
        const int k = j & (NS - 1);
        if constexpr (NS > 1) {
            float2 w = tw[k * (N2 / (NS * R))];
            if (INV) w.y = -w.y;
            v[1] = cmul(v[1], w);
            if constexpr (R >= 4) {
                float2 w2 = cmul(w, w), w3 = cmul(w2, w);
                v[2] = cmul(v[2], w2);
                v[3] = cmul(v[3], w3);
                if constexpr (R == 8) {
                    float2 w4 = cmul(w2, w2);
                    v[4] = cmul(v[4], w4);
                    v[5] = cmul(v[5], cmul(w4, w));
                    v[6] = cmul(v[6], cmul(w3, w3));
                    v[7] = cmul(v[7], cmul(w4, w3));
                }
            }
        }
        dft_r<R, INV>(v);
        if constexpr (TO_SMEM) {
            // j0 = (j / NS) * NS * R + k with j = l + G*b; G*b is a multiple of NS
            static_assert((Cfg::G % NS) == 0, "lane and butterfly fields must not overlap");
            const int lbase = fft_swz<N2>((l / NS) * NS * R + (l & (NS - 1)));
            const int cb = (Cfg::G * b / NS) * NS * R;
#pragma unroll
            for (int t = 0; t < R; ++t) buf[lbase ^ fft_swz<N2>(cb + t * NS)] = v[t];
        } else {
#pragma unroll
            for (int t = 0; t < R; ++t) X[b + t * NB] = v[t];
        }
    }
}

template <int N2>
__device__ __forceinline__ void fft_exchange(float2 (&X)[FftCfg<N2>::EPL], const float2 *buf, int l) {
    using Cfg = FftCfg<N2>;
    __syncwarp();
    const int lbase = fft_swz<N2>(l);
#pragma unroll
    for (int u = 0; u < Cfg::EPL; ++u) X[u] = buf[lbase ^ fft_swz<N2>(Cfg::G * u)];
    __syncwarp();
}

// In-register transform of the group's N2 points.  `buf` = this group's REGION of shared memory,
// `tw` = exp(-2*pi*i*q/N2), q < N2 (shared memory), l = lane index inside the group.
// All 32 lanes of the warp must call it together.
template <int N2, bool INV>
__device__ __forceinline__ void warp_fft(float2 (&X)[FftCfg<N2>::EPL], float2 *buf,
                                         const float2 *tw, int l) {
    using Cfg = FftCfg<N2>;
    fft_pass<N2, 8, 1, INV, true>(X, buf, tw, l);
    fft_exchange<N2>(X, buf, l);
    if constexpr (Cfg::R_LAST == 0) {
        fft_pass<N2, 8, 8, INV, false>(X, buf, tw, l);
    } else {
        fft_pass<N2, 8, 8, INV, true>(X, buf, tw, l);
        fft_exchange<N2>(X, buf, l);
        fft_pass<N2, Cfg::R_LAST, 64, INV, false>(X, buf, tw, l);
    }
}

}  // namespace rtb
