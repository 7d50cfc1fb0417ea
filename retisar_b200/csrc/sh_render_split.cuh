// K-B, experimental variant of the m-grouped render kernel (sh_render_mg.cuh) for 256-sample
// blocks: TWO warps per stream.  OFF by default (RTB_RENDER_SPLIT=1 enables it): measured on B200 it
// equals sh_render_mg_kernel at 1 024 streams (41.4 vs 41.0 us) and loses at 10 240 (298 vs 266 us)
// -- the named barriers and the extra shared-memory traffic eat what the doubled warp count gains
// (ncu: barrier 0.79, mio-throttle 0.62 stalls per issue; profiles/r02 notes).  Kept as the
// measured negative result of the "more warps per stream" direction.  Each thread owns 8 complex elements of the 512-point transforms
// (one radix-8 butterfly per pass) and half of the bins, so the per-thread state halves (about
// 100 registers instead of 168), twice as many warps are resident, and the critical path of one
// stream-block -- which sets the batch latency when the batch does not fill the GPU -- halves.
// The exchanges between the two warps go through the stream's shared-memory buffer and a named
// barrier (bar.sync id, 64); everything else (signal list, H tables, m-group flush, accumulators
// in shared memory) is the scheme of sh_render_mg.cuh.
#pragma once

#ifndef RTB_SPLIT_STREAMS
#define RTB_SPLIT_STREAMS 2   // streams per CTA (64 threads each)
#endif
#ifndef RTB_SPLIT_MIN_CTAS
#define RTB_SPLIT_MIN_CTAS 4
#endif

namespace rtb {

struct FftSplit512 {  // 512 points over 64 lanes
    static constexpr int LEN = 512, G = 64, EPL = 8, NPASS8 = 3, R_LAST = 0, REGION = 512;
};

__device__ __forceinline__ void named_barrier_64(int id) {
    asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
}

template <int NST>
__global__ void __launch_bounds__(64 * NST, RTB_SPLIT_MIN_CTAS)
sh_render_split_kernel(const ShRenderMgParams prm) {
    using Cfg = FftSplit512;
    constexpr int N2 = 512, B = 256, F = B + 1, G = 64, EPL = 8, HB = 4;
    constexpr int HSIG = 2 * F;
    constexpr int NT = 64 * NST;

    __shared__ float2 s_tw[N2];
    __shared__ __align__(128) float2 s_buf[NST][Cfg::REGION];
    __shared__ __align__(16) float4 s_h[2][HSIG];
    __shared__ float2 s_cs[NST][2][RTB_MAX_ORDER + 1];
    __shared__ ShMgSignal s_sig[RTB_MG_MAX_SIGNALS];  // the signal list: no dependent global load per round
    __shared__ __align__(16) float2 s_acc[NST][HB * 4 * G];  // [bin slot][cur L, cur R, last L, last R][lane]

    const int grp = threadIdx.x >> 6, l = threadIdx.x & 63;
    for (int i = threadIdx.x; i < N2; i += NT) s_tw[i] = prm.tw[i];
    for (int i = threadIdx.x; i < prm.np * (int)(sizeof(ShMgSignal) / sizeof(int2)); i += NT)
        reinterpret_cast<int2 *>(s_sig)[i] = __ldg(reinterpret_cast<const int2 *>(prm.signals) + i);
    __syncthreads();
    float2 *sacc = s_acc[grp] + l;
#pragma unroll
    for (int i = 0; i < HB * 4; ++i) sacc[i * G] = make_float2(0.f, 0.f);

    const int s = blockIdx.x * NST + grp;
    const bool active = s < prm.S;
    float2 *buf = s_buf[grp];
    float *zrow = reinterpret_cast<float *>(buf);
    const int np = prm.np;
    const size_t zoff = (size_t)(active ? s : 0) * prm.J * B;
    auto gsync = [&] { named_barrier_64(1 + grp); };

    auto prefetch_h = [&](int p, int stage) {
        const float4 *src = prm.htab + (size_t)p * HSIG;
        for (int i = threadIdx.x; i < HSIG; i += NT) cp_async16_ca(&s_h[stage][i], src + i);
    };
    auto prefetch_z = [&](int p) {
        if (!active) return;
        const int row_re = s_sig[p].row_re, row_im = s_sig[p].row_im;
        constexpr int CH = B / 4;  // 64 chunks of 16 bytes per half row
#pragma unroll
        for (int part = 0; part < 4; ++part) {  // re-prev, re-cur, im-prev, im-cur: one chunk per lane
            const int off = l * 4;
            const int row = (part < 2) ? row_re : row_im;
            float *dst = zrow + (part >> 1) * N2 + (part & 1) * B + off;
            if (row >= 0) {
                const float *src = ((part & 1) ? prm.zcur : prm.zprev) + zoff + (size_t)row * B + off;
                cp_async16_cg(dst, src);
            } else {
                *reinterpret_cast<float4 *>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        static_assert(CH == G, "one 16-byte chunk per lane and row half");
    };
    if (np > 0) {
        prefetch_h(0, 0);
        prefetch_z(0);
    }
    cp_async_commit();

    if (active) {  // a4: (cos, sin)(m alpha) for the current and the previous yaw
        const float rad_cur = yaw_deg_to_rad16(prm.yaw_deg[s], prm.tracker_dir, prm.src_azim16);
        const float rad_last = prm.rad_last_in[s];
        if (l <= prm.order_max) {
            float sn, cs;
            sincosf((float)l * rad_cur, &sn, &cs);
            s_cs[grp][0][l] = make_float2(cs, sn);
            sincosf((float)l * rad_last, &sn, &cs);
            s_cs[grp][1][l] = prm.last_on ? make_float2(cs, sn) : make_float2(0.f, 0.f);
        }
        if (prm.crossfade && l == 0) prm.rad_last_out[s] = rad_cur;
    }

    // group sums: U on bins l + 64u (slot HB = Nyquist, lane 0), V mirrored (slot v <-> element
    // HB + v, i.e. bin N2 - f; slot HB = bin 0, lane 0); Nyquist slot of the ear spectra in registers
    float2 U[HB + 1][2], V[HB + 1][2], accn[4];
#pragma unroll
    for (int u = 0; u <= HB; ++u) U[u][0] = U[u][1] = V[u][0] = V[u][1] = make_float2(0.f, 0.f);
#pragma unroll
    for (int q = 0; q < 4; ++q) accn[q] = make_float2(0.f, 0.f);

    int stage = 0;
    for (int p = 0; p < np; ++p, stage ^= 1) {
        cp_async_wait_all();
        __syncthreads();  // tile + rows have landed; everyone left the previous round
        if (p + 1 < np) prefetch_h(p + 1, stage ^ 1);
        cp_async_commit();
        if (!active) continue;  // uniform per stream group (named barriers are per group)
        const ShMgSignal sg = s_sig[p];

        float2 X[EPL];
#pragma unroll
        for (int u = 0; u < EPL; ++u) {
            X[u].x = zrow[l + G * u];
            X[u].y = zrow[N2 + l + G * u];
        }
        gsync();
        fft_run_c<Cfg, false>(X, buf, s_tw, l, gsync, [&] {
            if (!sg.flush) {  // the flush below still needs the buffer for its mirror exchange
                if (p + 1 < np) prefetch_z(p + 1);
                cp_async_commit();
            }
        });

        const float4 *h1 = s_h[stage], *h2 = h1 + F;
#pragma unroll
        for (int u = 0; u <= HB; ++u) {
            const float4 h = h1[(u < HB) ? (l + G * u) : B];
            cfma(U[u][0], make_float2(h.x, h.y), X[u]);
            cfma(U[u][1], make_float2(h.z, h.w), X[u]);
        }
        if (sg.has2) {
#pragma unroll
            for (int v = 0; v <= HB; ++v) {
                const float4 h = h2[(v < HB) ? (l + G * v) : B];
                const float2 x = (v < HB) ? X[HB + v] : X[0];
                cfma_conj(V[v][0], make_float2(h.x, h.y), x);
                cfma_conj(V[v][1], make_float2(h.z, h.w), x);
            }
        }
        if (sg.flush) {  // rotate the group sums by the two yaws and add them to the ear spectra
            const float2 p1c = phase_of(s_cs[grp][0], sg.m1), p1l = phase_of(s_cs[grp][1], sg.m1);
            const bool hasV = sg.m2 != RTB_MG_NO_TERM;
            float2 p2c = make_float2(0.f, 0.f), p2l = p2c;
            if (hasV) {
                p2c = phase_of(s_cs[grp][0], sg.m2);
                p2l = phase_of(s_cs[grp][1], sg.m2);
                // V at bin f lives at mirrored position B - f (lane (G-l)%G, slot HB-1-u): exchange
                // through the buffer, [ear][position]
#pragma unroll
                for (int v = 0; v < HB; ++v) {
                    buf[l + G * v] = V[v][0];
                    buf[B + l + G * v] = V[v][1];
                }
                gsync();
            }
#pragma unroll
            for (int u = 0; u <= HB; ++u) {
                float2 v0 = make_float2(0.f, 0.f), v1 = v0;
                if (hasV) {
                    if (u < HB) {
                        const int f = l + G * u;
                        if (u == 0 && l == 0) { v0 = V[HB][0]; v1 = V[HB][1]; }  // bin 0
                        else { v0 = buf[B - f]; v1 = buf[2 * B - f]; }
                    } else {
                        v0 = V[0][0]; v1 = V[0][1];  // Nyquist (lane 0)
                    }
                }
                float2 a[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) a[q] = (u < HB) ? sacc[(u * 4 + q) * G] : accn[q];
                cfma(a[0], p1c, U[u][0]);
                cfma(a[1], p1c, U[u][1]);
                cfma(a[2], p1l, U[u][0]);
                cfma(a[3], p1l, U[u][1]);
                if (hasV) {
                    cfma(a[0], p2c, v0);
                    cfma(a[1], p2c, v1);
                    cfma(a[2], p2l, v0);
                    cfma(a[3], p2l, v1);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (u < HB) sacc[(u * 4 + q) * G] = a[q];
                    else accn[q] = a[q];
                }
                U[u][0] = U[u][1] = make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u <= HB; ++u) V[u][0] = V[u][1] = make_float2(0.f, 0.f);
            if (hasV) gsync();  // mirror reads done: the buffer may receive the next rows
            if (p + 1 < np) prefetch_z(p + 1);
            cp_async_commit();
        }
    }
    cp_async_wait_all();
    if (np == 0) __syncthreads();
    if (!active) return;

    // ---- a6, a7: Hermitian extension of (L + iR), inverse transforms for both yaws, discard,
    // cross-fade.  Element N2 - f comes from lane (G-l)%G through the buffer.
    float2 ycur[HB], ylast[HB];
    const int niter = prm.crossfade ? 2 : 1;
    for (int it = 0; it < niter; ++it) {
        float2 X[EPL];
        gsync();
#pragma unroll
        for (int u = 0; u < HB; ++u) {
            float2 eL = sacc[(u * 4 + 2 * it) * G], eR = sacc[(u * 4 + 2 * it + 1) * G];
            if (u == 0 && l == 0) { eL.y = 0.f; eR.y = 0.f; }  // irfft ignores Im of the DC bin
            X[u] = make_float2(eL.x - eR.y, eL.y + eR.x);      // Z(f)    = L + i R
            buf[l + G * u] = make_float2(eL.x + eR.y, eR.x - eL.y);  // Z(N2-f) = conj(L) + i conj(R)
        }
        gsync();
#pragma unroll
        for (int u2 = HB; u2 < EPL; ++u2) {
            if (u2 == HB && l == 0) X[u2] = make_float2(accn[2 * it].x, accn[2 * it + 1].x);
            else X[u2] = buf[N2 - (l + G * u2)];
        }
        gsync();
        fft_run_c<Cfg, true>(X, buf, s_tw, l, gsync, [] {});
#pragma unroll
        for (int u = 0; u < HB; ++u) {
            if (it == 0) ycur[u] = X[u + HB];
            else ylast[u] = X[u + HB];
        }
    }
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

}  // namespace rtb
