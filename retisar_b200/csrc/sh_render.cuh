// K-B: the render kernel of the SH chain.  Included by sh_kernels.cuh (needs ShRenderParams).
//
// One warp renders one stream; the 4 warps of a CTA walk the signal list in lock step so that the
// HRIR tile of the signal(s) in flight is staged once per CTA in shared memory (double buffered,
// cp.async), and every warp prefetches the overlap-save rows of its next signal into its FFT
// exchange buffer while it is busy with the contraction of the current one.
#pragma once

#ifndef RTB_RENDER_MIN_CTAS
#define RTB_RENDER_MIN_CTAS 3
#endif

namespace rtb {

// cp.async helpers (LDGSTS): global -> shared without staging registers
__device__ __forceinline__ void cp_async16_cg(void *smem_dst, const void *gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16_ca(void *smem_dst, const void *gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// a6, a7: ear spectra (cur/last x left/right on the bins the lane owns) -> Hermitian extension,
// inverse FFT of (left + i right), overlap-save discard, cos^2 cross-fade, coalesced store.
// With G < 32 the current-yaw pair is transformed by group 0 and the previous-yaw pair by group 1
// at the same time; with G == 32 the warp does them one after the other.
// SOLO: the transform group does both yaws itself, one after the other, whatever the number of
// groups per warp (each group of the warp renders its own stream; `store` = the stream exists).
template <int N2, bool SOLO = false>
__device__ __forceinline__ void render_tail(float2 (&acc)[FftCfg<N2>::EPL / 2 + 1][4], float2 *buf,
                                            const float2 *s_tw, int grp, int l, const float *win_in,
                                            const float *win_out, int crossfade, float *o, bool store = true) {
    using Cfg = FftCfg<N2>;
    constexpr int G = Cfg::G, EPL = Cfg::EPL, HB = EPL / 2, GROUPS = SOLO ? 1 : Cfg::GROUPS;
    constexpr int B = N2 / 2;
    constexpr unsigned FULL = 0xffffffffu;
    fft_sync<N2>();
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
        fft_hermitian_fill<N2>(X, V, nyq, buf, l);
        warp_fft<N2, true>(X, buf, s_tw, l);
#pragma unroll
        for (int u = 0; u < HB; ++u) {
            if (GROUPS >= 2 || it == 0) ycur[u] = X[u + HB];
            if (GROUPS >= 2 || it == 1) ylast[u] = X[u + HB];
        }
        if (NITER == 2 && it == 0 && !crossfade) break;  // the previous-yaw block is not needed
    }
    if (GROUPS >= 2) {  // bring the previous-yaw block from group 1 to group 0
        const int lane = grp * G + l;
#pragma unroll
        for (int u = 0; u < HB; ++u) {
            ylast[u].x = __shfl_sync(FULL, ylast[u].x, (lane + G) & 31);
            ylast[u].y = __shfl_sync(FULL, ylast[u].y, (lane + G) & 31);
        }
    }
    if (grp == 0 && store) {
#pragma unroll
        for (int u = 0; u < HB; ++u) {
            const int n = l + G * u;
            float yl, yr;
            if (crossfade) {
                const float wi = __ldg(win_in + n), wo = __ldg(win_out + n);
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

template <int N2>
__global__ void __launch_bounds__(FftCfg<N2>::THREADS, FftCfg<N2>::BIG ? 1 : RTB_RENDER_MIN_CTAS)
sh_render_kernel(const ShRenderParams prm) {
    using Cfg = FftCfg<N2>;
    constexpr int G = Cfg::G, EPL = Cfg::EPL, HB = EPL / 2, GROUPS = Cfg::GROUPS;
    constexpr int B = N2 / 2, F = B + 1;
    constexpr int HSIG = 2 * F;  // float4 entries of one signal's HRIR tile (two terms)
    constexpr bool BIG = Cfg::BIG;  // one stream per CTA; HRIR tiles and twiddles straight from L2
    constexpr int NW = BIG ? 1 : 4;  // streams per CTA
    constexpr unsigned FULL = 0xffffffffu;

    extern __shared__ __align__(128) unsigned char dyn_smem[];
    __shared__ float2 s_tw[BIG ? 1 : N2];
    __shared__ __align__(128) float2 s_buf[BIG ? 1 : 4 * GROUPS * Cfg::REGION];
    __shared__ __align__(16) float4 s_h[2][BIG ? 1 : GROUPS * HSIG];
    __shared__ float2 s_cs[NW][2][RTB_MAX_ORDER + 1];

    const int warp = BIG ? 0 : (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int grp = BIG ? 0 : lane / G, l = BIG ? (int)threadIdx.x : lane % G;
    const float2 *tw = prm.tw;
    float2 *buf = reinterpret_cast<float2 *>(dyn_smem);
    if constexpr (!BIG) {
        for (int i = threadIdx.x; i < N2; i += 128) s_tw[i] = prm.tw[i];
        tw = s_tw;
        buf = s_buf + (warp * GROUPS + grp) * Cfg::REGION;
    }

    const int s = BIG ? (int)blockIdx.x : (int)blockIdx.x * 4 + warp;
    const bool active = s < prm.S;
    float *zrow = reinterpret_cast<float *>(buf);  // landing zone: re row [N2] | im row [N2]
    const int np = max(prm.np_cur, prm.np_last);
    const size_t zoff = (size_t)(active ? s : 0) * prm.J * B;

    // stage the HRIR tiles of signals p0 .. p0+GROUPS-1 (contiguous in the table)
    auto prefetch_h = [&](int p0, int stage) {
        if constexpr (!BIG) {
            const int n = min(GROUPS, np - p0) * HSIG;
            const float4 *src = prm.htab + (size_t)p0 * HSIG;
            for (int i = threadIdx.x; i < n; i += 128) cp_async16_ca(&s_h[stage][i], src + i);
        }
    };
    // bring the (previous | current) block of the group's signal into its exchange buffer
    auto prefetch_z = [&](int p0) {
        const int p = p0 + grp;
        if (!active || p >= np) return;
        const ShSignal sg = prm.signals[p];
        constexpr int CH = B / 4;  // 16-byte chunks per half row
        for (int i = l; i < 4 * CH; i += G) {
            const int part = i / CH, off = (i % CH) * 4;  // part: re-prev, re-cur, im-prev, im-cur
            const int row = (part < 2) ? sg.row_re : sg.row_im;
            float *dst = zrow + (part >> 1) * N2 + (part & 1) * B + off;
            if (row >= 0) {
                const float *src = ((part & 1) ? prm.zcur : prm.zprev) + zoff + (size_t)row * B + off;
                cp_async16_cg(dst, src);
            } else {
                *reinterpret_cast<float4 *>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    };
    if (np > 0) {
        prefetch_h(0, 0);
        prefetch_z(0);
    }
    cp_async_commit();

    // rotation phases for the current and the previous yaw (a4)
    if (active) {
        const float rad_cur = yaw_deg_to_rad16(prm.yaw_deg[s], prm.tracker_dir, prm.src_azim16);
        const float rad_last = prm.rad_last_in[s];
        if (threadIdx.x < 32 * NW && lane <= prm.order_max) {
            float sn, cs;
            sincosf((float)lane * rad_cur, &sn, &cs);
            s_cs[warp][0][lane] = make_float2(cs, sn);
            sincosf((float)lane * rad_last, &sn, &cs);
            s_cs[warp][1][lane] = make_float2(cs, sn);
        }
        if (prm.crossfade && lane == 0 && (!BIG || threadIdx.x == 0)) prm.rad_last_out[s] = rad_cur;
    }

    // accumulators: [bin slot][cur ear0, cur ear1, last ear0, last ear1]
    float2 acc[HB + 1][4];
#pragma unroll
    for (int u = 0; u <= HB; ++u)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[u][q] = make_float2(0.f, 0.f);

    int stage = 0;
    for (int p0 = 0; p0 < np; p0 += GROUPS, stage ^= 1) {
        cp_async_wait_all();
        __syncthreads();  // tile + rows of this round have landed; everyone left the last round
        if (p0 + GROUPS < np) prefetch_h(p0 + GROUPS, stage ^ 1);
        cp_async_commit();
        if (!BIG && !active) continue;

        const int p = p0 + grp;
        const bool valid = p < np;
        const ShSignal sg = prm.signals[valid ? p : 0];
        // ---- a1: overlap-save buffer (previous | current block) of the complex signal
        float2 X[EPL];
#pragma unroll
        for (int u = 0; u < EPL; ++u) {
            X[u].x = valid ? zrow[l + G * u] : 0.f;
            X[u].y = valid ? zrow[N2 + l + G * u] : 0.f;
        }
        fft_sync<N2>();
        // the exchange buffer is idle after the last exchange: prefetch the next signal's rows
        // (a CTA-wide transform still needs it for the mirror exchange below)
        warp_fft_hook<N2, false>(X, buf, tw, l, [&] {
            if constexpr (!BIG) {
                if (p0 + GROUPS < np) prefetch_z(p0 + GROUPS);
                cp_async_commit();
            }
        });
        float2 M[HB + 1];  // element N2-f for the bins this lane owns
        fft_mirror<N2>(X, M, buf, l);
        if constexpr (BIG) {
            if (p0 + GROUPS < np) prefetch_z(p0 + GROUPS);
            cp_async_commit();
        }

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
        const float4 *h1 = BIG ? prm.htab + (size_t)(valid ? p : 0) * HSIG
                               : s_h[stage] + (valid ? grp : 0) * HSIG;
        const float4 *h2 = h1 + F;
#pragma unroll
        for (int u = 0; u <= HB; ++u) {
            const int f = (u < HB) ? (l + G * u) : B;  // slot HB is the Nyquist bin (lane 0 only)
            const float2 zf = X[u];
            {
                const float4 h = BIG ? __ldg(h1 + f) : h1[f];
                const float2 tc = cmul(ph1c, zf), tl = cmul(ph1l, zf);
                cfma(acc[u][0], make_float2(h.x, h.y), tc);
                cfma(acc[u][1], make_float2(h.z, h.w), tc);
                cfma(acc[u][2], make_float2(h.x, h.y), tl);
                cfma(acc[u][3], make_float2(h.z, h.w), tl);
            }
            if (has2) {  // uniform inside a group
                const float4 h = BIG ? __ldg(h2 + f) : h2[f];
                const float2 zc = make_float2(M[u].x, -M[u].y);
                const float2 tc = cmul(ph2c, zc), tl = cmul(ph2l, zc);
                cfma(acc[u][0], make_float2(h.x, h.y), tc);
                cfma(acc[u][1], make_float2(h.z, h.w), tc);
                cfma(acc[u][2], make_float2(h.x, h.y), tl);
                cfma(acc[u][3], make_float2(h.z, h.w), tl);
            }
        }
    }
    cp_async_wait_all();
    if (np == 0) __syncthreads();  // twiddle table visibility when no round ran
    if (!active) return;

    // transforms that ran side by side in one warp each hold a partial sum
    if constexpr (!BIG) {
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
    }
    render_tail<N2>(acc, buf, tw, grp, l, prm.win_in, prm.win_out, prm.crossfade,
                    prm.out + (size_t)s * 2 * B);
    if (prm.hp.P > 0) hpcf_stage<N2>(prm.hp, s, grp == 0, buf, tw, l);  // prm.out is the HPCF's input state then
}

}  // namespace rtb
