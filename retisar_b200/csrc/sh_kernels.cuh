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
#include "rtb_async.cuh"

namespace rtb {

// ---------------------------------------------------------------------------------------------
// K-A: z[j][t] = sum_c Rm[j][c] x[c][t], the HBM-bound kernel of the chain.
//
// Persistent CTAs stream the input tiles [C][TW] of their streams into a ring of shared-memory
// stages with cp.async.bulk (TMA bulk copy, completion counted on an mbarrier), so that several
// tiles are in flight per SM regardless of occupancy.  Each warp then computes a (4*JTH rows) x
// (64 samples) tile with an SGEMM-style register micro-tile: lane = (row group jr = lane/8,
// column group tc = lane%8) owns JTH rows x 8 samples (two float4 runs, 32 samples apart, so that
// both the shared-memory reads and the global stores of a quarter warp are contiguous 128 bytes).
// The analysis matrix chunk is read as [C][32] (each row group padded to 8) from shared memory
// when all chunks fit, otherwise through the read-only L1 path.
// ---------------------------------------------------------------------------------------------
#define RTB_AN_MAX_STAGES 4

template <int JTH, int NRUN, bool R_SMEM>
__global__ void __launch_bounds__(128)
sh_analysis_kernel(const float *__restrict__ x,    // [S][C][B]
                   const float *__restrict__ rmP,  // [n_jchunks][C][32] packed analysis matrix
                   float *__restrict__ zcur,       // [S][J][B]
                   int S, int C, int J, int B, int TW, int n_jchunks, int stages) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[RTB_AN_MAX_STAGES], empty_bar[RTB_AN_MAX_STAGES];
    float *tiles = reinterpret_cast<float *>(smem_raw);
    const int tile_elems = C * TW;
    float *rsm = tiles + (size_t)stages * tile_elems;  // [n_jchunks][C][32] when R_SMEM
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int jr = lane >> 3, tc = lane & 7;
    const int n_wtiles = B / TW;  // tiles per stream
    const long long n_tiles_total = (long long)S * n_wtiles;
    const long long my_tiles = (n_tiles_total - blockIdx.x + gridDim.x - 1) / gridDim.x;
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (R_SMEM) {
        const int n4 = n_jchunks * C * 8;
        for (int i = threadIdx.x; i < n4; i += 128)
            reinterpret_cast<float4 *>(rsm)[i] = __ldg(reinterpret_cast<const float4 *>(rmP) + i);
    }
    __syncthreads();
    // PDL: the setup above overlapped the tail of the previous kernel in the stream; its output
    // (the other half of the z ring is still being read by the previous block's render kernel)
    // and the caller's input block are touched only from here on
    pdl_wait();
    pdl_trigger();

    auto issue = [&](long long it) {  // warp 0: bulk copies of tile `it` into its stage
        const long long tile = blockIdx.x + it * gridDim.x;
        const int st = (int)(it % stages);
        const long long s = tile / n_wtiles;
        const int t0 = (int)(tile % n_wtiles) * TW;
        if (lane == 0) mbar_expect_tx(&full_bar[st], (unsigned)(tile_elems * sizeof(float)));
        __syncwarp();
        const float *src = x + ((size_t)s * C) * B + t0;
        float *dst = tiles + (size_t)st * tile_elems;
        if (TW == B) {  // the whole [C][B] block of the stream is contiguous
            if (lane == 0) bulk_copy_g2s(dst, src, (unsigned)(tile_elems * sizeof(float)), &full_bar[st]);
        } else {
            for (int c = lane; c < C; c += 32)
                bulk_copy_g2s(dst + (size_t)c * TW, src + (size_t)c * B, (unsigned)(TW * sizeof(float)), &full_bar[st]);
        }
    };
    if (warp == 0)
        for (int i = 0; i < stages && i < my_tiles; ++i) issue(i);

    constexpr int WT = 32 * NRUN, NX = 4 * NRUN;  // samples per warp tile / per lane
    const int n_tsub = TW / WT;
    const int n_tasks = n_tsub * n_jchunks;
    for (long long it = 0; it < my_tiles; ++it) {
        const int st = (int)(it % stages);
        mbar_wait(&full_bar[st], (unsigned)((it / stages) & 1));
        const long long tile = blockIdx.x + it * gridDim.x;
        const long long s = tile / n_wtiles;
        const int tbase = (int)(tile % n_wtiles) * TW;
        const float *xs = tiles + (size_t)st * tile_elems;
        for (int task = warp; task < n_tasks; task += 4) {
            const int jc = task / n_tsub, ts = task % n_tsub;
            const int tl = ts * WT + tc * 4;  // first run; the second starts 32 samples later
            const float *rp = (R_SMEM ? rsm : rmP) + (size_t)jc * C * 32 + jr * 8;
            // accumulators as (sample, sample + 1) pairs: one FFMA2 with the matrix element
            // broadcast does two MACs per issue slot
            f32x2_t acc[JTH][NX / 2];
#pragma unroll
            for (int j = 0; j < JTH; ++j)
#pragma unroll
                for (int i = 0; i < NX / 2; ++i) acc[j][i] = pk2(0.f, 0.f);
#pragma unroll 4
            for (int c = 0; c < C; ++c) {
                f32x2_t xv[NX / 2];
#pragma unroll
                for (int q = 0; q < NRUN; ++q) {
                    const float4 xa = *reinterpret_cast<const float4 *>(xs + (size_t)c * TW + tl + 32 * q);
                    xv[2 * q] = pk2(xa.x, xa.y);
                    xv[2 * q + 1] = pk2(xa.z, xa.w);
                }
                float r[8];
                const float4 *rq = reinterpret_cast<const float4 *>(rp + (size_t)c * 32);
                const float4 r0 = R_SMEM ? rq[0] : __ldg(rq);
                r[0] = r0.x; r[1] = r0.y; r[2] = r0.z; r[3] = r0.w;
                if (JTH > 4) {
                    const float4 r1 = R_SMEM ? rq[1] : __ldg(rq + 1);
                    r[4] = r1.x; r[5] = r1.y; r[6] = r1.z; r[7] = r1.w;
                }
#pragma unroll
                for (int j = 0; j < JTH; ++j)
#pragma unroll
                    for (int i = 0; i < NX / 2; ++i) acc[j][i] = fma2(xv[i], pk2(r[j], r[j]), acc[j][i]);
            }
            const int jbase = jc * 4 * JTH + jr * JTH;
#pragma unroll
            for (int j = 0; j < JTH; ++j) {
                if (jbase + j < J) {
                    float *zp = zcur + ((size_t)s * J + jbase + j) * B + tbase + tl;
#pragma unroll
                    for (int q = 0; q < NRUN; ++q) {
                        const float2 lo = up2(acc[j][2 * q]), hi = up2(acc[j][2 * q + 1]);
                        *reinterpret_cast<float4 *>(zp + 32 * q) = make_float4(lo.x, lo.y, hi.x, hi.y);
                    }
                }
            }
        }
        // release the stage (one arrival per warp); warp 0 refills it once all four have left,
        // the other warps run ahead into the next stage
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[st]);
        if (warp == 0 && it + stages < my_tiles) {
            mbar_wait(&empty_bar[st], (unsigned)((it / stages) & 1));
            issue(it + stages);
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

}  // namespace rtb
#include "hpcf_stage.cuh"
namespace rtb {

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
    HpcfParams hp;               // optional HPCF stage behind the cross-fade (hp.P == 0: none)
};

// phase = exp(-i * m * alpha) from the (cos, sin)(|m| alpha) table
__device__ __forceinline__ float2 phase_of(const float2 *cs, int m) {
    float2 v = cs[m < 0 ? -m : m];
    return make_float2(v.x, m < 0 ? v.y : -v.y);
}

}  // namespace rtb

#include "sh_analysis_tc.cuh"
#include "sh_analysis_mma.cuh"
#include "sh_render.cuh"
#include "sh_render_mg.cuh"
#include "sh_partitioned.cuh"
#include "sh_partitioned_tc.cuh"

namespace rtb {

// a10: passthrough.  irfft(rfft(buffer))[B:] is the current block itself, so the first E input
// channels are copied to the ears (_convolver.py:558-562); missing channels stay silent.
static __global__ void sh_passthrough_kernel(const float *__restrict__ x, float *__restrict__ out,
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
