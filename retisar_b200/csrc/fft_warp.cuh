// Warp-level complex FFT (Stockham autosort, radix 8/4/2) held in registers, with the inter-pass
// exchange through a warp-private, XOR-swizzled shared-memory buffer.
//
// A transform of N2 points is owned by a "group" of G lanes (G = 32 for N2 = 256 and 512, N2/8
// below, so small transforms run 2 or 4 per warp side by side; N2/16 lanes = a whole CTA for
// N2 >= 1024, with the same code and __syncthreads between the passes).  Lane l of a group holds EPL = N2/G
// complex values X[u] <-> element index l + G*u, before the first and after the last pass.  That
// layout is what the callers build directly from coalesced global loads and consume directly in
// registers (no shared-memory round trip for the spectrum): elements u < EPL/2 of the result are
// the bins 0 <= f < N2/2 owned by the lane, and element N2-f lives in lane (G-l)%G.
//
// Convention matches numpy: forward unnormalised exp(-2*pi*i*n*k/N2), inverse exp(+...) also
// unnormalised (the caller folds 1/N2 into its window).
#pragma once
#include "rtb_common.cuh"

#ifndef RTB_FFT_TWTAB
// twiddle powers by complex multiplication (0) or from the table (1).  Measured on B200 (render
// kernel, 10 240 streams): table 282 us vs multiplication 251 us -- the strided table reads cost
// more shared-memory wavefronts than the 12 packed FP instructions per butterfly they replace.
#define RTB_FFT_TWTAB 0
#endif

namespace rtb {

template <int N2>
struct FftCfg {
    static_assert(N2 >= 64 && N2 <= 8192 && (N2 & (N2 - 1)) == 0, "unsupported FFT length");
    // lanes per transform: a fraction of a warp (N2 < 256), one warp, or -- for N2 >= 1024 -- a
    // whole CTA of N2/16 threads ("big" transforms: exchanges go through __syncthreads)
    static constexpr int LEN = N2;
    static constexpr int G = (N2 >= 512) ? N2 / 16 : ((N2 >= 256) ? 32 : N2 / 8);
    static constexpr bool BIG = G > 32;
    static constexpr int EPL = N2 / G;                   // complex elements per lane (8 or 16)
    static constexpr int GROUPS = BIG ? 1 : 32 / G;      // transforms per warp
    // radix-8 passes through shared memory, then one register pass of radix R_LAST (0 = none)
    static constexpr int NPASS8 = (N2 >= 4096) ? 4 : ((N2 >= 512) ? 3 : 2);
    static constexpr int R_LAST = N2 >> (3 * NPASS8) == 1 ? 0 : N2 >> (3 * NPASS8);
    static constexpr int REGION = N2 + (N2 == 64 ? 8 : 0);  // float2 slots per transform buffer
    static constexpr int THREADS = BIG ? G : 128;        // CTA size of the kernels built on it
    static constexpr int TASKS = BIG ? 1 : 4 * GROUPS;   // transforms side by side per CTA
};

// barrier over the lanes of one transform
template <int N2>
__device__ __forceinline__ void fft_sync() {
    if constexpr (FftCfg<N2>::BIG) __syncthreads();
    else __syncwarp();
}

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

// Shared-memory accesses by 32-bit shared address.  The swizzle only changes index bits 0..3, and
// the lane part and the compile-time part of every index occupy disjoint bit fields above them, so
//   address(lane ^ const) = ((base + 8 * lane) ^ (8 * const & 0x78)) + (8 * const & ~0x78)
// for any exchange buffer aligned to 128 bytes: one LOP3 per distinct low part (a handful per
// pass) and an immediate offset per access, instead of XOR + scale + add for each of them.
// (RTB_FFT_XORADDR = 0 restores plain indexing; transforms of 64 points always use it.)
#ifndef RTB_FFT_XORADDR
#define RTB_FFT_XORADDR 1
#endif
__device__ __forceinline__ unsigned smem_addr32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sts64(unsigned addr, float2 v) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ float2 lds64(unsigned addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
    return v;
}

// multiply by -i (forward) / +i (inverse)
template <bool INV>
__device__ __forceinline__ float2 rot90(float2 a) {
    return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}

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
    // d1 = e1 * exp(-+i pi/4) = h (e1 + rot90(e1)),  d3 = e3 * exp(-+3i pi/4) = h (rot90(e3) - e3)
    float2 d1 = cscale(cadd(e1, rot90<INV>(e1)), h);
    float2 d3 = cscale(csub(rot90<INV>(e3), e3), h);
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
// Cfg: LEN (transform length), G (lanes per transform), EPL (elements per lane).
template <class Cfg, int R, int NS, bool INV, bool TO_SMEM, bool AL = (RTB_FFT_XORADDR && Cfg::LEN >= 128)>
__device__ __forceinline__ void fft_pass_c(float2 (&X)[Cfg::EPL], float2 *buf, const float2 *tw, int l) {
    constexpr int N2 = Cfg::LEN;
    constexpr int NB = Cfg::EPL / R;  // butterflies per lane
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        float2 v[R];
#pragma unroll
        for (int t = 0; t < R; ++t) v[t] = X[b + t * NB];
        const int j = l + Cfg::G * b;
        const int k = j & (NS - 1);
        if constexpr (NS > 1) {
#if RTB_FFT_TWTAB
            // all powers w^t straight from the table (it holds exp(-2 pi i q / N2) for every q):
            // R-1 loads instead of one load and R-2 dependent complex multiplies
            constexpr int STEP = N2 / (NS * R);
#pragma unroll
            for (int t = 1; t < R; ++t) {
                float2 w = tw[k * (t * STEP)];
                if (INV) w.y = -w.y;
                v[t] = cmul(v[t], w);
            }
#else
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
#endif
        }
        dft_r<R, INV>(v);
        if constexpr (TO_SMEM) {
            // j0 = (j / NS) * NS * R + k with j = l + G*b; G*b is a multiple of NS
            static_assert((Cfg::G % NS) == 0, "lane and butterfly fields must not overlap");
            const int lbase = fft_swz<N2>((l / NS) * NS * R + (l & (NS - 1)));
            const int cb = (Cfg::G * b / NS) * NS * R;
            if constexpr (AL) {
                const unsigned pre = smem_addr32(buf) + (unsigned)lbase * 8u;
#pragma unroll
                for (int t = 0; t < R; ++t) {
                    const unsigned c8 = (unsigned)fft_swz<N2>(cb + t * NS) * 8u;
                    sts64((pre ^ (c8 & 0x78u)) + (c8 & ~0x78u), v[t]);
                }
            } else {
#pragma unroll
                for (int t = 0; t < R; ++t) buf[lbase ^ fft_swz<N2>(cb + t * NS)] = v[t];
            }
        } else {
#pragma unroll
            for (int t = 0; t < R; ++t) X[b + t * NB] = v[t];
        }
    }
}

// gather the next pass's inputs from the exchange buffer; `sync` = barrier over the transform's lanes
template <class Cfg, bool AL = (RTB_FFT_XORADDR && Cfg::LEN >= 128), class Sync>
__device__ __forceinline__ void fft_exchange_c(float2 (&X)[Cfg::EPL], const float2 *buf, int l, Sync &&sync) {
    sync();
    const int lbase = fft_swz<Cfg::LEN>(l);
    if constexpr (AL) {
        const unsigned pre = smem_addr32(buf) + (unsigned)lbase * 8u;
#pragma unroll
        for (int u = 0; u < Cfg::EPL; ++u) {
            const unsigned c8 = (unsigned)fft_swz<Cfg::LEN>(Cfg::G * u) * 8u;
            X[u] = lds64((pre ^ (c8 & 0x78u)) + (c8 & ~0x78u));
        }
    } else {
#pragma unroll
        for (int u = 0; u < Cfg::EPL; ++u) X[u] = buf[lbase ^ fft_swz<Cfg::LEN>(Cfg::G * u)];
    }
    sync();
}

// ---------------------------------------------------------------------------------------------
// 512 points on one warp as 16 x 16 x 2 (RTB_FFT_R16): the kernels built on the warp transform are
// bound by the L1 / shared-memory data pipe (render kernel: 78 % of its peak, 42 % of the wavefronts
// are the two exchanges of the 8 x 8 x 8 schedule).  With radix 16 the lane's 16 elements are one
// butterfly, so only ONE exchange goes through shared memory; the second one, in front of the final
// radix-2 pass, is a swap of eight values with lane l ^ 16 (shuffles).  Same Stockham index algebra
// as fft_pass_c: pass (R, NS) reads X[t] <-> element l + 32 t, multiplies by w^t with
// w = exp(-+2 pi i k / (NS R)), k = l % NS, and leaves output t at position (l / NS) NS R + k + t NS.
// ---------------------------------------------------------------------------------------------
// Measured (render kernel, c2 10 240 streams / c5 8 192 streams): 230 vs 240 us, 517 vs 539 us.
#ifndef RTB_FFT_R16
#define RTB_FFT_R16 1
#endif

// exchange-buffer placement of the 16 x 16 schedule: writes 16 l + t (l = lane) and reads l + 32 u
// are both conflict free per half warp (8-byte accesses)
__host__ __device__ __forceinline__ constexpr int fft_swz16(int i) { return i ^ ((i >> 4) & 15); }

template <bool INV>
__device__ __forceinline__ float2 cmul_w16(float2 a, int p) {  // a * exp(-+2 pi i p / 16), p compile time
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
    float2 w;
    switch (p & 15) {
        case 0: return a;
        case 1: w = make_float2(c1, -s1); break;
        case 2: w = make_float2(h, -h); break;
        case 3: w = make_float2(s1, -c1); break;
        case 4: return rot90<INV>(a);
        case 5: w = make_float2(-s1, -c1); break;
        case 6: w = make_float2(-h, -h); break;
        case 7: w = make_float2(-c1, -s1); break;
        case 8: return make_float2(-a.x, -a.y);
        case 9: w = make_float2(-c1, s1); break;
        case 10: w = make_float2(-h, h); break;
        case 11: w = make_float2(-s1, c1); break;
        case 12: { float2 r = rot90<INV>(a); return make_float2(-r.x, -r.y); }
        case 13: w = make_float2(s1, c1); break;
        case 14: w = make_float2(h, h); break;
        default: w = make_float2(c1, s1); break;
    }
    if (INV) w.y = -w.y;
    return cmul(a, w);
}

// 16-point DFT in place (4 x 4 decomposition): natural order in, output t in slot dft16_out(t)
template <bool INV>
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) dft4<INV>(v[q], v[q + 4], v[q + 8], v[q + 12]);  // v[q + 4 r] = a[q][r]
#pragma unroll
    for (int q = 1; q < 4; ++q)
#pragma unroll
        for (int r = 1; r < 4; ++r) v[q + 4 * r] = cmul_w16<INV>(v[q + 4 * r], q * r);
#pragma unroll
    for (int r = 0; r < 4; ++r) dft4<INV>(v[4 * r], v[4 * r + 1], v[4 * r + 2], v[4 * r + 3]);  // over q -> X[r + 4 s]
    // v[4 r + s] now holds output r + 4 s: the callers index the result through dft16_out()
}
// register slot of output t of dft16
__host__ __device__ __forceinline__ constexpr int dft16_out(int t) { return 4 * (t & 3) + (t >> 2); }

template <bool INV, class Sync, class Hook>
__device__ __forceinline__ void fft512_r16(float2 (&X)[16], float2 *buf, const float2 *tw, int l,
                                           Sync &&sync, Hook &&after_last_exchange) {
    // pass A: radix 16, NS = 1 (no twiddles); output t of lane l -> position 16 l + t
    dft16<INV>(X);
    // addresses as (lane part) ^ (immediate) + (immediate): the buffer is 128-byte aligned, the lane
    // parts and the compile-time parts occupy disjoint bit fields
    {
        const unsigned pre = smem_addr32(buf) + 128u * (unsigned)l + 8u * (unsigned)(l & 15);
#pragma unroll
        for (int t = 0; t < 16; ++t) sts64(pre ^ (8u * t), X[dft16_out(t)]);  // index 16 l + (t ^ l % 16)
    }
    sync();
    {
        const unsigned hi = (unsigned)l >> 4;
        const unsigned pre = smem_addr32(buf) + 8u * (16u * hi + (((unsigned)l & 15u) ^ hi));
#pragma unroll
        for (int u = 0; u < 16; ++u)  // index (l + 32 u) ^ ((l / 16 + 2 u) % 16)
            X[u] = lds64((pre ^ (8u * ((2 * u) & 15))) + 256u * u);
    }
    sync();
    after_last_exchange();
    // pass B: radix 16, NS = 16: k = l % 16, w = exp(-+2 pi i k / 256) = tw[2 k]
    {
        float2 w = tw[2 * (l & 15)];
        if (INV) w.y = -w.y;
        float2 wp = w;
#pragma unroll
        for (int t = 1; t < 16; ++t) {
            X[t] = cmul(X[t], wp);
            if (t < 15) wp = cmul(wp, w);
        }
    }
    dft16<INV>(X);
    // output t of lane l sits at position p = (l / 16) 256 + l % 16 + 16 t.  The final radix-2 pass
    // (NS = 256) of lane l needs elements j = l + 32 b and j + 256, b < 8: both have t = l / 16 + 2 b
    // and live in lanes l % 16 and l % 16 + 16 -- one of them is this lane, the other lane l ^ 16.
    const bool upper = l >= 16;
    float2 wl = tw[l];  // exp(-2 pi i l / 512); times exp(-2 pi i b / 16) per butterfly
    if (INV) wl.y = -wl.y;
    float2 Y[16];
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const float2 even = X[dft16_out(2 * b)], odd = X[dft16_out(2 * b + 1)];
        const float2 send = upper ? even : odd;
        float2 recv;
        recv.x = __shfl_xor_sync(0xffffffffu, send.x, 16);
        recv.y = __shfl_xor_sync(0xffffffffu, send.y, 16);
        float2 v0 = upper ? recv : even;
        float2 v1 = upper ? odd : recv;
        v1 = cmul(v1, cmul_w16<INV>(wl, b));
        Y[b] = cadd(v0, v1);
        Y[b + 8] = csub(v0, v1);
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) X[u] = Y[u];
}

// The same idea for 128 points on half a warp (64-sample blocks: 16 lanes x 8 elements) as 8 x 8 x 2:
// one exchange through shared memory (placement i ^ (i / 8) % 16: writes 8 l + t and reads l + 16 u
// are conflict free per group), then a swap of four values with lane l ^ 8 and the radix-2 pass.
#ifndef RTB_FFT_R8X2
#define RTB_FFT_R8X2 1
#endif
template <bool INV, class Sync, class Hook>
__device__ __forceinline__ void fft128_r8(float2 (&X)[8], float2 *buf, const float2 *tw, int l,
                                          Sync &&sync, Hook &&after_last_exchange) {
    dft8<INV>(X);  // pass A: radix 8, NS = 1; output t of lane l -> position 8 l + t
    {
        const unsigned pre = smem_addr32(buf) + 8u * (unsigned)((8 * l) ^ l);
#pragma unroll
        for (int t = 0; t < 8; ++t) sts64(pre ^ (8u * t), X[t]);
    }
    sync();
    {
        const unsigned pre = smem_addr32(buf) + 8u * (unsigned)(l ^ (l >> 3));
#pragma unroll
        for (int u = 0; u < 8; ++u) X[u] = lds64((pre ^ (8u * ((2 * u) & 15))) + 128u * u);
    }
    sync();
    after_last_exchange();
    {   // pass B: radix 8, NS = 8: k = l % 8, w = exp(-+2 pi i k / 64) = tw[2 k]
        float2 w = tw[2 * (l & 7)];
        if (INV) w.y = -w.y;
        float2 wp = w;
#pragma unroll
        for (int t = 1; t < 8; ++t) {
            X[t] = cmul(X[t], wp);
            if (t < 7) wp = cmul(wp, w);
        }
    }
    dft8<INV>(X);
    // output t of lane l sits at position (l / 8) 64 + l % 8 + 8 t; the radix-2 pass (NS = 64) of lane l
    // needs elements l + 16 b and l + 16 b + 64, b < 4: t = l / 8 + 2 b in lanes l % 8 and l % 8 + 8
    const bool upper = l >= 8;
    float2 wl = tw[l];
    if (INV) wl.y = -wl.y;
    float2 Y[8];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const float2 even = X[2 * b], odd = X[2 * b + 1];
        const float2 send = upper ? even : odd;
        float2 recv;
        recv.x = __shfl_xor_sync(0xffffffffu, send.x, 8);
        recv.y = __shfl_xor_sync(0xffffffffu, send.y, 8);
        const float2 v0 = upper ? recv : even;
        float2 v1 = upper ? odd : recv;
        v1 = cmul(v1, cmul_w16<INV>(wl, 2 * b));
        Y[b] = cadd(v0, v1);
        Y[b + 4] = csub(v0, v1);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) X[u] = Y[u];
}

// In-register transform of the group's points.  `buf` = this group's REGION of shared memory,
// `tw` = exp(-2*pi*i*q/LEN), q < LEN (shared or global memory), l = lane index inside the group.
// All lanes of the group must call it together.  `after_last_exchange()` runs once the exchange
// buffer is idle (before the final register pass): the place to start prefetching the next input.
template <class Cfg, bool INV, bool AL = (RTB_FFT_XORADDR && Cfg::LEN >= 128), class Sync, class Hook>
__device__ __forceinline__ void fft_run_c(float2 (&X)[Cfg::EPL], float2 *buf, const float2 *tw, int l,
                                          Sync &&sync, Hook &&after_last_exchange) {
    constexpr int NP = Cfg::NPASS8;
    constexpr bool TAIL = Cfg::R_LAST != 0;
    if constexpr (RTB_FFT_R16 && Cfg::LEN == 512 && Cfg::G == 32 && Cfg::EPL == 16) {
        fft512_r16<INV>(X, buf, tw, l, sync, after_last_exchange);
        return;
    }
    if constexpr (RTB_FFT_R8X2 && Cfg::LEN == 128 && Cfg::G == 16 && Cfg::EPL == 8) {
        fft128_r8<INV>(X, buf, tw, l, sync, after_last_exchange);
        return;
    }
    fft_pass_c<Cfg, 8, 1, INV, true, AL>(X, buf, tw, l);
    fft_exchange_c<Cfg, AL>(X, buf, l, sync);
    if constexpr (NP == 2 && !TAIL) after_last_exchange();
    fft_pass_c<Cfg, 8, 8, INV, (NP > 2 || TAIL), AL>(X, buf, tw, l);
    if constexpr (NP > 2 || TAIL) fft_exchange_c<Cfg, AL>(X, buf, l, sync);
    if constexpr (NP >= 3) {
        if constexpr (NP == 3 && !TAIL) after_last_exchange();
        fft_pass_c<Cfg, 8, 64, INV, (NP > 3 || TAIL), AL>(X, buf, tw, l);
        if constexpr (NP > 3 || TAIL) fft_exchange_c<Cfg, AL>(X, buf, l, sync);
    }
    if constexpr (NP >= 4) {
        if constexpr (!TAIL) after_last_exchange();
        fft_pass_c<Cfg, 8, 512, INV, TAIL, AL>(X, buf, tw, l);
        if constexpr (TAIL) fft_exchange_c<Cfg, AL>(X, buf, l, sync);
    }
    if constexpr (TAIL) {
        after_last_exchange();
        fft_pass_c<Cfg, Cfg::R_LAST, (1 << (3 * NP)), INV, false>(X, buf, tw, l);
    }
}

template <int N2, bool INV, bool AL = (RTB_FFT_XORADDR && N2 >= 128), class Hook>
__device__ __forceinline__ void warp_fft_hook(float2 (&X)[FftCfg<N2>::EPL], float2 *buf,
                                              const float2 *tw, int l, Hook &&after_last_exchange) {
    fft_run_c<FftCfg<N2>, INV, AL>(X, buf, tw, l, [] { fft_sync<N2>(); }, after_last_exchange);
}

template <int N2, bool INV, bool AL = (RTB_FFT_XORADDR && N2 >= 128)>
__device__ __forceinline__ void warp_fft(float2 (&X)[FftCfg<N2>::EPL], float2 *buf,
                                         const float2 *tw, int l) {
    warp_fft_hook<N2, INV, AL>(X, buf, tw, l, [] {});
}

// After a forward transform lane l holds element l + G*u in X[u].  Mirror access: the element at
// index (N2 - f) mod N2 for the bins f = l + G*u, u < EPL/2, the lane owns, plus (slot EPL/2) the
// Nyquist element itself.  Inside a warp that is slot EPL-1-u of lane G-l (lane 0: its own
// slots); a CTA-wide transform goes through the exchange buffer (natural order).
template <int N2>
__device__ __forceinline__ void fft_mirror(const float2 (&X)[FftCfg<N2>::EPL],
                                           float2 (&M)[FftCfg<N2>::EPL / 2 + 1], float2 *buf, int l) {
    using Cfg = FftCfg<N2>;
    constexpr int G = Cfg::G, EPL = Cfg::EPL, HB = EPL / 2;
    if constexpr (Cfg::BIG) {
#pragma unroll
        for (int u = HB; u < EPL; ++u) buf[l + G * u] = X[u];
        __syncthreads();
#pragma unroll
        for (int u = 0; u < HB; ++u) M[u] = (u == 0 && l == 0) ? X[0] : buf[N2 - (l + G * u)];
        M[HB] = X[HB];
        __syncthreads();
    } else {
        const int partner = (G - l) & (G - 1);
#pragma unroll
        for (int u = 0; u < HB; ++u) {
            float2 zp;
            zp.x = __shfl_sync(0xffffffffu, X[EPL - 1 - u].x, partner, G);
            zp.y = __shfl_sync(0xffffffffu, X[EPL - 1 - u].y, partner, G);
            if (l == 0) zp = (u == 0) ? X[0] : X[(EPL - u) % EPL];
            M[u] = zp;
        }
        M[HB] = X[HB];
    }
}

// Hermitian extension before an inverse transform: V[u] is the value wanted at element index
// N2 - f for the bins f = l + G*u (u < EPL/2) this lane owns; fills X[EPL/2 .. EPL-1] (element
// l + G*u2), with `nyq` for element N2/2.
template <int N2>
__device__ __forceinline__ void fft_hermitian_fill(float2 (&X)[FftCfg<N2>::EPL],
                                                   const float2 (&V)[FftCfg<N2>::EPL / 2], float2 nyq,
                                                   float2 *buf, int l) {
    using Cfg = FftCfg<N2>;
    constexpr int G = Cfg::G, EPL = Cfg::EPL, HB = EPL / 2;
    if constexpr (Cfg::BIG) {
        __syncthreads();
#pragma unroll
        for (int u = 0; u < HB; ++u) buf[l + G * u] = V[u];
        __syncthreads();
#pragma unroll
        for (int u2 = HB; u2 < EPL; ++u2)
            X[u2] = (u2 == HB && l == 0) ? nyq : buf[N2 - (l + G * u2)];
        __syncthreads();
    } else {
        const int partner = (G - l) & (G - 1);
#pragma unroll
        for (int u2 = HB; u2 < EPL; ++u2) {
            float2 x;
            x.x = __shfl_sync(0xffffffffu, V[EPL - 1 - u2].x, partner, G);
            x.y = __shfl_sync(0xffffffffu, V[EPL - 1 - u2].y, partner, G);
            if (l == 0) x = (u2 == HB) ? nyq : V[(EPL - u2) % HB];
            X[u2] = x;
        }
    }
}

// Launch of a kernel built on FftCfg<N2>: `tasks` transforms-in-flight (streams, channel pairs...),
// TASKS per CTA; CTA-wide transforms get their exchange buffer as dynamic shared memory.
template <int N2, class Kernel, class Params, class... Extra>
inline void fft_kernel_launch(Kernel kernel, const Params &prm, long long tasks, cudaStream_t st,
                              size_t extra_smem = 0, int small_tasks_per_cta = 0, Extra... extra) {
    using Cfg = FftCfg<N2>;
    const size_t smem = (Cfg::BIG ? sizeof(float2) * Cfg::REGION : 0) + extra_smem;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int per_cta = Cfg::BIG ? 1 : (small_tasks_per_cta ? small_tasks_per_cta : Cfg::TASKS);
    const unsigned grid = (unsigned)((tasks + per_cta - 1) / per_cta);
    kernel<<<grid, Cfg::THREADS, smem, st>>>(prm, extra...);
}

// N2 = 2 * block length dispatch over every supported size
#define RTB_DISPATCH_N2(N2V, CALL)                                                                  \
    switch (N2V) {                                                                                  \
        case 64: CALL(64); break;     case 128: CALL(128); break;   case 256: CALL(256); break;     \
        case 512: CALL(512); break;   case 1024: CALL(1024); break; case 2048: CALL(2048); break;   \
        case 4096: CALL(4096); break; default: CALL(8192); break;                                   \
    }

inline bool block_length_supported(int b) { return b >= 32 && b <= 4096 && (b & (b - 1)) == 0; }
#define RTB_BLOCK_LENGTHS "32, 64, 128, 256, 512, 1024, 2048, 4096"

}  // namespace rtb
