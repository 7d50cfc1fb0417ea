// K-B, even/odd-bin variant of the m-grouped render kernel for 256-sample blocks.  OFF by default
// (RTB_RENDER_EO=1): parity-green, but measured on B200 it only ties sh_render_mg_kernel at
// 512 ... 2 048 streams (26.9 / 39.2 / 71.3 vs 29.2 / 38.6 / 71.7 us) and loses at 10 240 (284 vs
// 252 us).  A clock64 trace (tools/eo_trace.py) shows why: with 16 warps per SM the schedulers are
// issue-saturated, and the two half-chains execute 24 % more instruction cycles per stream than one
// full chain (both warps read all rows, pre/post twiddles, duplicated loop overhead).  The render
// kernels are bound by instruction throughput, not by the latency of one warp's chain.
//
// The 512-point transform of the overlap-save buffer [a | b] (a = previous, b = current block)
// splits by one decimation-in-frequency stage into two independent 256-point transforms:
//     even bins  Z[2j]   = FFT256(a + b)[j]
//     odd bins   Z[2j+1] = FFT256((a - b) * w^n)[j],   w = exp(-2 pi i / 512)
// The HRIR multiply, the rotation and the sum over SH coefficients are per bin, and the wanted
// second half of the inverse transform recombines in the time domain,
//     y[256 + n] = ( IFFT256(Y_even)[n] - conj(w^n) * IFFT256(Y_odd)[n] ) / 512,
// so the whole chain of one stream decomposes into two half-size chains that meet only in that last
// line.  Here they run on TWO WARPS of the same CTA that never synchronise inside the transforms
// (unlike sh_render_split_kernel): each keeps 8 complex values, 4 + 4 bin slots of the group sums
// and a 4 KB accumulator, about 100 registers -- 16 warps (8 streams) per SM instead of 12 warps
// (12 streams), and the critical path of one stream-block, which sets the latency of a batch that
// does not fill the GPU, halves.  Everything else is the scheme of sh_render_mg.cuh (signal list
// grouped by order m, V sums kept mirrored, real rows two per transform, accumulators in shared
// memory); the H tables are the grouped ones, decimated per role (sh_plan.cu: build_eo_tables).
#pragma once

#ifndef RTB_EO_STREAMS
#define RTB_EO_STREAMS 4   // streams per CTA (2 warps each)
#endif

namespace rtb {

// -DRTB_TRACE: warp 0 of CTA 0 writes clock64 stamps of its loop phases to a device buffer
#ifdef RTB_TRACE
__device__ long long g_eo_trace[1024];
#define EO_STAMP(i) do { if (blockIdx.x == RTB_TRACE && warp == 0 && lane == 0 && (i) < 1024) g_eo_trace[(i)] = clock64(); } while (0)
#else
#define EO_STAMP(i) do {} while (0)
#endif

#define RTB_EO_HE 129  // float4 entries per (role, term) of one signal's tile

template <int NST>
__global__ void __launch_bounds__(64 * NST, 2)
sh_render_eo_kernel(const ShRenderMgParams prm) {
    using Cfg = FftCfg<256>;
    constexpr int B = 256, NH = 256, HB = 4, EPL = 8;
    constexpr int HE = RTB_EO_HE, HSIG = 4 * HE;  // [role][term][HE]
    constexpr int NT = 64 * NST, NW = 2 * NST;
    constexpr unsigned FULL = 0xffffffffu;

    extern __shared__ __align__(128) unsigned char dyn_smem[];
    float2 *s_acc_all = reinterpret_cast<float2 *>(dyn_smem);                 // [NW][HB*4*32]
    float *s_land_all = reinterpret_cast<float *>(s_acc_all + NW * HB * 4 * 32);  // [NST][4][B] rows
    float2 *s_buf_all = reinterpret_cast<float2 *>(s_land_all + NST * 4 * B);     // [NW][NH] exchange
    __shared__ __align__(16) float2 s_tw512[B];   // w^n, n < 256
    __shared__ __align__(16) float2 s_tw256[NH];  // exp(-2 pi i q / 256)
    __shared__ __align__(16) float4 s_h[2][HSIG];
    __shared__ float2 s_cs[NST][2][RTB_MAX_ORDER + 1];
    __shared__ ShMgSignal s_sig[RTB_MG_MAX_SIGNALS];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    EO_STAMP(0);
    const int ss = warp >> 1;          // stream slot in the CTA
    const bool odd = warp & 1;         // role: even or odd bins
    for (int i = threadIdx.x; i < B / 2; i += NT) cp_async16_ca(&s_tw512[2 * i], prm.tw + 2 * i);
    for (int i = threadIdx.x; i < NH; i += NT) s_tw256[i] = __ldg(prm.tw + 2 * i);
    for (int i = threadIdx.x; i < prm.np * (int)(sizeof(ShMgSignal) / sizeof(int2)); i += NT)
        reinterpret_cast<int2 *>(s_sig)[i] = __ldg(reinterpret_cast<const int2 *>(prm.signals) + i);
    float2 *sacc = s_acc_all + warp * (HB * 4 * 32) + lane;
#pragma unroll
    for (int i = 0; i < HB * 4; ++i) sacc[i * 32] = make_float2(0.f, 0.f);
    __syncthreads();

    const int s = blockIdx.x * NST + ss;
    const bool active = s < prm.S;
    float2 *buf = s_buf_all + warp * NH;
    float *land = s_land_all + ss * 4 * B;  // [re prev | re cur | im prev | im cur]
    const int np = prm.np;
    const size_t zoff = (size_t)(active ? s : 0) * prm.J * B;
    auto pair_sync = [&] { asm volatile("bar.sync %0, 64;" ::"r"(1 + ss) : "memory"); };

    auto prefetch_h = [&](int p, int stage) {
        const float4 *src = prm.htab + (size_t)p * HSIG;
        for (int i = threadIdx.x; i < HSIG; i += NT) cp_async16_ca(&s_h[stage][i], src + i);
    };
    // the even warp fetches the real-part rows, the odd warp the imaginary-part rows
    auto prefetch_z = [&](int p) {
        if (!active) return;
        const int row = odd ? s_sig[p].row_im : s_sig[p].row_re;
        float *dst = land + (odd ? 2 * B : 0) + lane * 4;
        if (row >= 0) {
            const size_t o = zoff + (size_t)row * B + lane * 4;
#pragma unroll
            for (int c = 0; c < B / 128; ++c) {
                cp_async16_cg(dst + 128 * c, prm.zprev + o + 128 * c);
                cp_async16_cg(dst + B + 128 * c, prm.zcur + o + 128 * c);
            }
        } else {
#pragma unroll
            for (int c = 0; c < 2 * B / 128; ++c)
                *reinterpret_cast<float4 *>(dst + 128 * c) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    if (np > 0) {
        prefetch_h(0, 0);
        prefetch_z(0);
    }
    cp_async_commit();

    if (active && !odd) {  // a4: (cos, sin)(m alpha) for the current and the previous yaw
        const float rad_cur = yaw_deg_to_rad16(prm.yaw_deg[s], prm.tracker_dir, prm.src_azim16);
        const float rad_last = prm.rad_last_in[s];
        if (lane <= prm.order_max) {
            float sn, cs;
            sincosf((float)lane * rad_cur, &sn, &cs);
            s_cs[ss][0][lane] = make_float2(cs, sn);
            sincosf((float)lane * rad_last, &sn, &cs);
            s_cs[ss][1][lane] = prm.last_on ? make_float2(cs, sn) : make_float2(0.f, 0.f);
        }
        if (prm.crossfade && lane == 0) prm.rad_last_out[s] = rad_cur;
    }

    // Group sums.  Even role: U on elements j = l + 32u (bin 2j; slot HB = element 128 = Nyquist,
    // lane 0), V on elements 128 + l + 32v (slot HB = element 0 = DC, lane 0).  Odd role: U on
    // elements j (bin 2j+1), V on elements 128 + i (mirror of bin 255 - 2i); no special slots.
    float2 U[HB + 1][2], V[HB + 1][2], accn[4];
#pragma unroll
    for (int u = 0; u <= HB; ++u) U[u][0] = U[u][1] = V[u][0] = V[u][1] = make_float2(0.f, 0.f);
#pragma unroll
    for (int q = 0; q < 4; ++q) accn[q] = make_float2(0.f, 0.f);
    const int partner = odd ? 31 - lane : ((32 - lane) & 31);

    int stage = 0;
    for (int p = 0; p < np; ++p, stage ^= 1) {
        EO_STAMP(8 + 8 * p + 0);
        cp_async_wait_all();
        EO_STAMP(8 + 8 * p + 1);
        __syncthreads();  // tile + rows have landed; everyone left the previous round
        EO_STAMP(8 + 8 * p + 2);
        if (p + 1 < np) prefetch_h(p + 1, stage ^ 1);
        cp_async_commit();
        if (!active) continue;  // uniform per stream pair
        const ShMgSignal sg = s_sig[p];

        // ---- a1: decimation-in-frequency stage on the overlap-save buffer
        float2 X[EPL];
#pragma unroll
        for (int u = 0; u < EPL; ++u) {
            const int n = lane + 32 * u;
            const float2 a = make_float2(land[n], land[2 * B + n]);
            const float2 b = make_float2(land[B + n], land[3 * B + n]);
            X[u] = odd ? cmul(csub(a, b), s_tw512[n]) : cadd(a, b);
        }
        EO_STAMP(8 + 8 * p + 3);
        pair_sync();  // both roles have read the rows: the landing zone may receive the next ones
        if (p + 1 < np) prefetch_z(p + 1);
        cp_async_commit();
        EO_STAMP(8 + 8 * p + 4);
        fft_run_c<Cfg, false>(X, buf, s_tw256, lane, [] { __syncwarp(); }, [] {});
        EO_STAMP(8 + 8 * p + 5);

        // ---- a3: HRIR multiply into the group sums
        const float4 *h1 = s_h[stage] + (odd ? 2 * HE : 0), *h2 = h1 + HE;
#pragma unroll
        for (int u = 0; u < HB; ++u) {
            const float4 h = h1[lane + 32 * u];
            cfma(U[u][0], make_float2(h.x, h.y), X[u]);
            cfma(U[u][1], make_float2(h.z, h.w), X[u]);
        }
        if (!odd) {
            const float4 h = h1[128];
            cfma(U[HB][0], make_float2(h.x, h.y), X[HB]);
            cfma(U[HB][1], make_float2(h.z, h.w), X[HB]);
        }
        if (sg.has2) {
#pragma unroll
            for (int v = 0; v < HB; ++v) {
                const float4 h = h2[lane + 32 * v];
                cfma_conj(V[v][0], make_float2(h.x, h.y), X[HB + v]);
                cfma_conj(V[v][1], make_float2(h.z, h.w), X[HB + v]);
            }
            if (!odd) {
                const float4 h = h2[128];
                cfma_conj(V[HB][0], make_float2(h.x, h.y), X[0]);
                cfma_conj(V[HB][1], make_float2(h.z, h.w), X[0]);
            }
        }
        EO_STAMP(8 + 8 * p + 6);
        // ---- a4, a5: rotate the group sums by the two yaws and add them to the ear spectra
        if (sg.flush) {
            const float2 p1c = phase_of(s_cs[ss][0], sg.m1), p1l = phase_of(s_cs[ss][1], sg.m1);
            const bool hasV = sg.m2 != RTB_MG_NO_TERM;
            float2 p2c = make_float2(0.f, 0.f), p2l = p2c;
            if (hasV) {
                p2c = phase_of(s_cs[ss][0], sg.m2);
                p2l = phase_of(s_cs[ss][1], sg.m2);
            }
#pragma unroll
            for (int u = 0; u <= HB; ++u) {
                if (u == HB && odd) break;  // no special slot in the odd role
                float2 v0 = make_float2(0.f, 0.f), v1 = v0;
                if (hasV) {
                    if (u < HB) {  // mirrored slot HB-1-u of the partner lane
                        v0.x = __shfl_sync(FULL, V[HB - 1 - u][0].x, partner);
                        v0.y = __shfl_sync(FULL, V[HB - 1 - u][0].y, partner);
                        v1.x = __shfl_sync(FULL, V[HB - 1 - u][1].x, partner);
                        v1.y = __shfl_sync(FULL, V[HB - 1 - u][1].y, partner);
                        if (!odd && lane == 0) { v0 = V[HB - u][0]; v1 = V[HB - u][1]; }
                    } else {
                        v0 = V[0][0]; v1 = V[0][1];  // Nyquist (even role, lane 0)
                    }
                }
                float2 a[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) a[q] = (u < HB) ? sacc[(u * 4 + q) * 32] : accn[q];
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
                    if (u < HB) sacc[(u * 4 + q) * 32] = a[q];
                    else accn[q] = a[q];
                }
                U[u][0] = U[u][1] = make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u <= HB; ++u) V[u][0] = V[u][1] = make_float2(0.f, 0.f);
        }
    }
    EO_STAMP(1);
    cp_async_wait_all();
    __syncthreads();
    if (!active) return;

    // ---- a6, a7: per role one inverse 256-point transform per yaw; the odd role hands
    // conj(w^n) * IFFT(Y_odd) over through the (now idle) landing zone; the even role combines,
    // cross-fades and stores
    float2 *xch = reinterpret_cast<float2 *>(land);  // [2 yaws][256]
    float2 ye[2][EPL];
    const int niter = prm.crossfade ? 2 : 1;
    for (int it = 0; it < 2; ++it) {
        if (it >= niter) break;
        float2 X[EPL], Vh[HB];
#pragma unroll
        for (int u = 0; u < HB; ++u) {
            float2 eL = sacc[(u * 4 + 2 * it) * 32], eR = sacc[(u * 4 + 2 * it + 1) * 32];
            if (!odd && u == 0 && lane == 0) { eL.y = 0.f; eR.y = 0.f; }  // irfft ignores Im of the DC bin
            X[u] = make_float2(eL.x - eR.y, eL.y + eR.x);   // L + i R
            Vh[u] = make_float2(eL.x + eR.y, eR.x - eL.y);  // conj(L) + i conj(R)
        }
        if (!odd) {
            fft_hermitian_fill<256>(X, Vh, make_float2(accn[2 * it].x, accn[2 * it + 1].x), buf, lane);
        } else {
#pragma unroll
            for (int v = 0; v < HB; ++v) {  // element 128 + l + 32v mirrors bin 255 - 2(l + 32v)
                X[HB + v].x = __shfl_sync(FULL, Vh[HB - 1 - v].x, partner);
                X[HB + v].y = __shfl_sync(FULL, Vh[HB - 1 - v].y, partner);
            }
        }
        __syncwarp();
        fft_run_c<Cfg, true>(X, buf, s_tw256, lane, [] { __syncwarp(); }, [] {});
        if (odd) {
#pragma unroll
            for (int u = 0; u < EPL; ++u) {
                const int n = lane + 32 * u;
                const float2 w = s_tw512[n];
                xch[it * B + n] = cmul(X[u], make_float2(w.x, -w.y));
            }
        } else {
#pragma unroll
            for (int u = 0; u < EPL; ++u) ye[it][u] = X[u];
        }
    }
    EO_STAMP(2);
    pair_sync();
    if (odd) return;
    float *o = prm.out + (size_t)s * 2 * B;
#pragma unroll
    for (int u = 0; u < EPL; ++u) {
        const int n = lane + 32 * u;
        const float2 yc = csub(ye[0][u], xch[n]);
        float yl, yr;
        if (prm.crossfade) {
            const float2 yp = csub(ye[1][u], xch[B + n]);
            const float wi = __ldg(prm.win_in + n), wo = __ldg(prm.win_out + n);
            yl = yc.x * wi + yp.x * wo;
            yr = yc.y * wi + yp.y * wo;
        } else {
            const float sc = 1.0f / 512.0f;
            yl = yc.x * sc;
            yr = yc.y * sc;
        }
        o[n] = yl;
        o[B + n] = yr;
    }
    EO_STAMP(3);
}

}  // namespace rtb
