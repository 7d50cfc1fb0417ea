// K-B, steady-state variant: the render kernel with the SH-coefficient sum grouped by order m.
// Included by sh_kernels.cuh after sh_render.cuh (shares its cp.async helpers and render_tail).
//
// The head rotation multiplies every coefficient (n, m) by exp(-i m alpha): all signals of one m
// share their two phases.  So instead of rotating every (signal, term) product for the current and
// the previous yaw (sh_render_kernel: 2 complex multiplies + 4 complex MACs per term and bin), the
// yaw-independent sums
//     U_m(ear, f) = sum_n H1[n,m](ear, f) * Z_nm(f)
//     V_m(ear, f) = sum_n H2[n,m](ear, f) * conj(Z_nm(N2 - f))
// are accumulated first (2 complex MACs per term and bin), and the rotation
//     ear_cur += e^{-i m1 a_cur} U_m + e^{-i m2 a_cur} V_m     (same for the previous yaw)
// is applied once per group.  V_m is accumulated where bin N2-f lives after the FFT (upper half of
// the lane's elements, H2 stored mirrored) and brought to bin f by one shuffle per group instead of
// one per signal.  The real (m = 0) analysis rows are transformed two per complex FFT: with
// Z = A + iB the sum H_a A + H_b B is again a two-term signal with H1 = (H_a - i H_b)/2,
// H2 = (H_a + i H_b)/2 and both phases 1.
//
// Valid while the current and the previous yaw use the same SH order (always, except for the one
// block after update_sh_processing); the plan falls back to sh_render_kernel otherwise.
#pragma once

// Ear-spectrum accumulators: they are touched only once per m group, so they live in shared memory
// (8 KB per stream) and the kernel keeps the register budget of three CTAs per SM; 0 = registers
// (255 registers, two CTAs per SM).
#ifndef RTB_MG_ACC_SMEM
#define RTB_MG_ACC_SMEM 1
#endif
#ifndef RTB_RENDER_MG_MIN_CTAS
#define RTB_RENDER_MG_MIN_CTAS (RTB_MG_ACC_SMEM ? 3 : 2)
#endif

namespace rtb {

#define RTB_MG_MAX_SIGNALS 96  // staged in shared memory (order 12 has 85)

#define RTB_MG_NO_TERM 127

struct __align__(8) ShMgSignal {  // one complex FFT signal of the grouped list (sh_plan.cu: build_mg_tables)
    short row_re;        // analysis row giving the real part
    short row_im;        // analysis row giving the imaginary part, -1 = zero
    signed char m1;      // group phase of the U sum: exp(-i*m1*alpha)
    signed char m2;      // group phase of the V sum, RTB_MG_NO_TERM = the group has no V sum
    unsigned char has2;  // this signal contributes to V
    unsigned char flush; // last signal of its group
};

struct ShRenderMgParams {
    const float *zprev, *zcur;   // [S][J][B]
    const float4 *htab;          // [NPm][2][F]: H1[f], then H2 mirrored (index i <-> bin B - i)
    const ShMgSignal *signals;   // [NPm]
    const float2 *tw;
    const float *win_in, *win_out;
    const float *yaw_deg, *rad_last_in;
    float *rad_last_out;
    float *out;
    int S, J, B;
    int np;                      // signals
    int last_on;                 // previous-yaw sum active (0: first block / cross-fade off)
    int order_max;
    int crossfade;
    float tracker_dir, src_azim16;
    HpcfParams hp;               // optional HPCF stage behind the cross-fade (hp.P == 0: none)
};

// acc += h * conj(x)
__device__ __forceinline__ void cfma_conj(float2 &acc, float2 h, float2 x) {
#if RTB_PACKED_F32
    acc = up2(fma2(pk2(h.y, -h.x), pk2(x.y, x.y), fma2(pk2(h), pk2(x.x, x.x), pk2(acc))));
#else
    acc.x = fmaf(h.x, x.x, acc.x);
    acc.x = fmaf(h.y, x.y, acc.x);
    acc.y = fmaf(h.y, x.x, acc.y);
    acc.y = fmaf(-h.x, x.y, acc.y);
#endif
}

// One transform group (G lanes: a whole warp for N2 = 256 / 512, half a warp for N2 = 128, a quarter
// for N2 = 64) renders one stream; the groups of a warp are different streams walking the signal
// list in lock step, so a CTA covers 4 * GROUPS streams and shares the staged HRIR tile among them.
// NW warps per CTA: 4 (NW * GROUPS streams share the staged HRIR tile), or 2 for small batches of
// short blocks, where 8 - 16 streams per CTA would leave SMs idle.
template <int N2, int NW = 4>
__global__ void __launch_bounds__(32 * NW, RTB_RENDER_MG_MIN_CTAS * (4 / NW))
sh_render_mg_kernel(const ShRenderMgParams prm) {
    using Cfg = FftCfg<N2>;
    static_assert(!Cfg::BIG, "warp-level transforms only");
    constexpr int G = Cfg::G, GROUPS = Cfg::GROUPS, EPL = Cfg::EPL, HB = EPL / 2;
    constexpr int B = N2 / 2, F = B + 1;
    constexpr int HSIG = 2 * F;
    constexpr unsigned FULL = 0xffffffffu;

    __shared__ __align__(16) float2 s_tw[N2];
    constexpr int NT = 32 * NW;
    __shared__ __align__(128) float2 s_buf[NW][GROUPS * Cfg::REGION];
    __shared__ __align__(16) float4 s_h[2][HSIG];
    __shared__ float2 s_cs[NW * GROUPS][2][RTB_MAX_ORDER + 1];
    __shared__ ShMgSignal s_sig[RTB_MG_MAX_SIGNALS];  // the signal list: no dependent global load per round
#if RTB_MG_ACC_SMEM
    extern __shared__ __align__(128) unsigned char dyn_smem[];  // [4 warps][HB][4][32 lanes] float2
#endif

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane / G, l = lane % G;        // transform group of the lane, lane inside the group
    const int sw = warp * GROUPS + grp;            // stream slot inside the CTA
    // twiddles arrive asynchronously with the first tile (waited for at the top of the first round)
    for (int i = threadIdx.x; i < N2 / 2; i += NT) cp_async16_ca(&s_tw[2 * i], prm.tw + 2 * i);
    for (int i = threadIdx.x; i < prm.np * (int)(sizeof(ShMgSignal) / sizeof(int2)); i += NT)
        reinterpret_cast<int2 *>(s_sig)[i] = __ldg(reinterpret_cast<const int2 *>(prm.signals) + i);
    __syncthreads();
#if RTB_MG_ACC_SMEM
    float2 *sacc = reinterpret_cast<float2 *>(dyn_smem) + warp * (HB * 4 * 32) + lane;
#pragma unroll
    for (int i = 0; i < HB * 4; ++i) sacc[i * 32] = make_float2(0.f, 0.f);
#endif

    const int s = blockIdx.x * (NW * GROUPS) + sw;
    const bool active = s < prm.S;
    float2 *buf = s_buf[warp] + grp * Cfg::REGION;
    float *zrow = reinterpret_cast<float *>(buf);
    const int partner = (G - l) & (G - 1);
    const int np = prm.np;
    const size_t zoff = (size_t)(active ? s : 0) * prm.J * B;

    auto prefetch_h = [&](int p, int stage) {
        const float4 *src = prm.htab + (size_t)p * HSIG + threadIdx.x;
        float4 *dst = &s_h[stage][threadIdx.x];
#pragma unroll
        for (int i = 0; i < HSIG / NT; ++i) cp_async16_ca(dst + NT * i, src + NT * i);
        if (threadIdx.x < HSIG % NT) cp_async16_ca(dst + (HSIG / NT) * NT, src + (HSIG / NT) * NT);
    };
    // rows of the signal -> landing zone [re: prev | cur][im: prev | cur]; lane-contiguous 16-byte chunks
    auto prefetch_z = [&](int p) {
        if (!active) return;
        const int row_re = s_sig[p].row_re, row_im = s_sig[p].row_im;
        constexpr int NCH = B / (4 * G);  // 16-byte chunks per lane and row half
        static_assert(NCH >= 1, "a row half has at least one chunk per lane");
#pragma unroll
        for (int im = 0; im < 2; ++im) {
            const int row = im ? row_im : row_re;
            float *dst = zrow + im * N2 + l * 4;
            if (row >= 0) {
                const size_t o = zoff + (size_t)row * B + l * 4;
                const float *sp = prm.zprev + o, *sc = prm.zcur + o;
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    cp_async16_cg(dst + 4 * G * c, sp + 4 * G * c);
                    cp_async16_cg(dst + B + 4 * G * c, sc + 4 * G * c);
                }
            } else {
#pragma unroll
                for (int c = 0; c < 2 * NCH; ++c)
                    *reinterpret_cast<float4 *>(dst + 4 * G * c) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    };
    if (np > 0) prefetch_h(0, 0);
    // PDL: twiddles, signal list and the first HRIR tile were requested while the analysis kernel
    // was still draining; its rows (and the caller's yaw angles) are read only after the wait
    pdl_wait();
    pdl_trigger();
    if (np > 0) prefetch_z(0);
    cp_async_commit();

    if (active) {  // a4: (cos, sin)(m alpha) for the current and the previous yaw
        const float rad_cur = yaw_deg_to_rad16(prm.yaw_deg[s], prm.tracker_dir, prm.src_azim16);
        const float rad_last = prm.rad_last_in[s];
        for (int m = l; m <= prm.order_max; m += G) {
            float sn, cs;
            sincosf((float)m * rad_cur, &sn, &cs);
            s_cs[sw][0][m] = make_float2(cs, sn);
            sincosf((float)m * rad_last, &sn, &cs);
            s_cs[sw][1][m] = prm.last_on ? make_float2(cs, sn) : make_float2(0.f, 0.f);
        }
        if (prm.crossfade && l == 0) prm.rad_last_out[s] = rad_cur;
    }

    // ear spectra [bin slot][cur ear0, cur ear1, last ear0, last ear1]; group sums U (bins
    // l + 32u, slot HB = Nyquist on lane 0) and V (mirrored: slot v <-> element HB + v, i.e. bin
    // N2 - f; slot HB = bin 0 on lane 0)
    float2 U[HB + 1][2], V[HB + 1][2];
#if RTB_MG_ACC_SMEM
    float2 accn[4];  // Nyquist slot (lane 0)
#pragma unroll
    for (int q = 0; q < 4; ++q) accn[q] = make_float2(0.f, 0.f);
#else
    float2 acc[HB + 1][4];
#pragma unroll
    for (int u = 0; u <= HB; ++u)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[u][q] = make_float2(0.f, 0.f);
#endif
#pragma unroll
    for (int u = 0; u <= HB; ++u) U[u][0] = U[u][1] = V[u][0] = V[u][1] = make_float2(0.f, 0.f);

    int stage = 0;
    for (int p = 0; p < np; ++p, stage ^= 1) {
        cp_async_wait_all();
        __syncthreads();
        if (p + 1 < np) prefetch_h(p + 1, stage ^ 1);
        cp_async_commit();
        if (GROUPS == 1 && !active) continue;   // (with several streams per warp the whole warp stays in step)
        const ShMgSignal sg = s_sig[p];

        float2 X[EPL];
#pragma unroll
        for (int u = 0; u < EPL; ++u) {
            X[u].x = active ? zrow[l + G * u] : 0.f;
            X[u].y = active ? zrow[N2 + l + G * u] : 0.f;
        }
        __syncwarp();
        warp_fft_hook<N2, false>(X, buf, s_tw, l, [&] {
            if (p + 1 < np) prefetch_z(p + 1);  // the exchange buffer is idle from here on
            cp_async_commit();
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
            const float2 p1c = phase_of(s_cs[sw][0], sg.m1), p1l = phase_of(s_cs[sw][1], sg.m1);
            const bool hasV = sg.m2 != RTB_MG_NO_TERM;
            float2 p2c = make_float2(0.f, 0.f), p2l = p2c;
            if (hasV) {
                p2c = phase_of(s_cs[sw][0], sg.m2);
                p2l = phase_of(s_cs[sw][1], sg.m2);
            }
#pragma unroll
            for (int u = 0; u <= HB; ++u) {
                // V at bin f = lane + 32u: mirrored slot HB-1-u of the partner lane; lane 0
                // keeps its own slots (bin 0 <-> slot HB, bin 32u <-> slot HB-u, Nyquist <-> 0)
                float2 v0 = make_float2(0.f, 0.f), v1 = v0;
                if (hasV) {
                    if (u < HB) {
                        v0.x = __shfl_sync(FULL, V[HB - 1 - u][0].x, partner, G);
                        v0.y = __shfl_sync(FULL, V[HB - 1 - u][0].y, partner, G);
                        v1.x = __shfl_sync(FULL, V[HB - 1 - u][1].x, partner, G);
                        v1.y = __shfl_sync(FULL, V[HB - 1 - u][1].y, partner, G);
                        if (l == 0) { v0 = V[HB - u][0]; v1 = V[HB - u][1]; }
                    } else {
                        v0 = V[0][0]; v1 = V[0][1];
                    }
                }
                float2 a[4];
#if RTB_MG_ACC_SMEM
#pragma unroll
                for (int q = 0; q < 4; ++q) a[q] = (u < HB) ? sacc[(u * 4 + q) * 32] : accn[q];
#else
#pragma unroll
                for (int q = 0; q < 4; ++q) a[q] = acc[u][q];
#endif
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
#if RTB_MG_ACC_SMEM
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (u < HB) sacc[(u * 4 + q) * 32] = a[q];
                    else accn[q] = a[q];
                }
#else
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[u][q] = a[q];
#endif
                U[u][0] = U[u][1] = make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u <= HB; ++u) V[u][0] = V[u][1] = make_float2(0.f, 0.f);
        }
    }
    cp_async_wait_all();
    if (np == 0) __syncthreads();
    if (GROUPS == 1 && !active) return;
    if (!__any_sync(FULL, active)) return;
#if RTB_MG_ACC_SMEM
    float2 acc[HB + 1][4];
#pragma unroll
    for (int u = 0; u <= HB; ++u)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[u][q] = (u < HB) ? sacc[(u * 4 + q) * 32] : accn[q];
#endif
    // every group renders its own stream: both yaws one after the other inside the group
    render_tail<N2, true>(acc, buf, s_tw, 0, l, prm.win_in, prm.win_out, prm.crossfade,
                          prm.out + (size_t)(active ? s : 0) * 2 * B, active);
    if (prm.hp.P > 0) hpcf_stage<N2>(prm.hp, active ? s : 0, active, buf, s_tw, l);  // prm.out is the HPCF's input state then
}

}  // namespace rtb
