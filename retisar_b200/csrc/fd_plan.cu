// HRIR / BRIR switching renderer for S batched streams (reference: AdjustableFdConvolver,
// ReTiSAR/_convolver.py:672-975, with a FilterSetSsr, _filter_set.py:843-899).
//
// Per block and stream: every source signal is transformed (overlap-save), the filter pair of the
// direction int(-azimuth) is picked per source from the head yaw (float16 chain, :947-975), and
// H[dir][p] * X / n_sources is added into slot p of an OUTPUT-side queue of P ear spectra
// (:838-866); a second queue receives the previous block's directions; the heads are inverse
// transformed, cross-faded (:833-836) and leave the queue (:645-665).  The filter changes from
// block to block, so unlike the time-invariant OLS convolver the partition sum has to stay on the
// output side.  One warp per stream, spectra and the queue head in registers.
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <new>
#include <vector>

#include "sh_kernels.cuh"

using namespace rtb;

namespace {

struct FdParams {
    const float *x;            // [S][NS][B]
    const float *xprev;        // [S][NS][B] state (read)
    float *xnext;              // [S][NS][B] state (write)
    const float2 *hfd;         // [n_dirs][P][2][Fp]
    float2 *qcur, *qlast;      // [S][P][2][Fp] output-side queues (rings)
    const float *yaw_deg;      // [S]
    const float *src_azim16;   // [NS] source azimuths, already rounded to float16
    const int *dir_last_in;    // [S][NS] direction index of the previous block
    int *dir_last_out;         // [S][NS]
    const float2 *tw;
    const float *win_in, *win_out;
    float *out;                // [S][2][B]
    float2 *spec;              // [S][NS][Fp] scaled source spectra (split launch: head kernel -> queue kernel)
    int *dir_cur;              // [S][NS] direction index of this block (split launch)
    int S, NS, P, Fp, n_dirs, head_cur, head_last, crossfade, last_valid, passthrough;
    float tracker_dir, inv_ns;
};

// float16 azimuth chain of _calculate_individual_directions -> direction index int(-az)
__device__ __forceinline__ int direction_index(float yaw_deg, float tracker_dir, float src16, int n_dirs) {
    float a = rn16(tracker_dir * yaw_deg);
    a = rn16(a + src16);
    a = rn16(a + 360.0f);
    float r = fmodf(a, 360.0f);
    if (r != 0.0f && r < 0.0f) r += 360.0f;
    if (!(r >= 0.0f)) return 0;  // NaN yaw (|yaw| beyond float16): the reference fails here
    const int k = (int)r;  // int(-az) == -trunc(az); python negative indexing wraps once
    return k == 0 ? 0 : n_dirs - k;
}

// PB = partitions whose loads are requested together, MINB = resident CTAs per SM the register
// allocation is limited to (two shapes are launched, see fd_dense below)
// HEAD_ONLY: only partition 0 (the queue heads, which stay in registers) is computed here and the
// scaled source spectra + direction indices are left for fd_queue_kernel
template <int N2, int PB, int MINB, bool HEAD_ONLY = false>
__global__ void __launch_bounds__(FftCfg<N2>::THREADS, MINB)
fd_render_kernel(const FdParams prm) {
    using Cfg = FftCfg<N2>;
    constexpr int G = Cfg::G, EPL = Cfg::EPL, HB = EPL / 2, GROUPS = Cfg::GROUPS, B = N2 / 2;
    constexpr bool BIG = Cfg::BIG;  // one stream per CTA, exchange buffer in dynamic smem
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
    const int s = BIG ? (int)blockIdx.x : (int)blockIdx.x * 4 + warp;
    if (s >= prm.S) return;
    // with G < 32 every group redundantly processes the same stream (sources are few); group 0 writes
    const int Fp = prm.Fp, NS = prm.NS;
    const float yaw = prm.passthrough ? 0.f : prm.yaw_deg[s];

    float2 acc[HB + 1][4];  // queue heads: [bin slot][cur L, cur R, last L, last R]
    float2 *qc = prm.qcur + (size_t)s * prm.P * 2 * Fp, *ql = prm.qlast + (size_t)s * prm.P * 2 * Fp;
    const bool do_last = prm.crossfade && !prm.passthrough;
#pragma unroll
    for (int u = 0; u <= HB; ++u) {
        const int f = (u < HB) ? (l + G * u) : B;
        acc[u][0] = qc[((size_t)prm.head_cur * 2 + 0) * Fp + f];
        acc[u][1] = qc[((size_t)prm.head_cur * 2 + 1) * Fp + f];
        acc[u][2] = do_last ? ql[((size_t)prm.head_last * 2 + 0) * Fp + f] : make_float2(0.f, 0.f);
        acc[u][3] = do_last ? ql[((size_t)prm.head_last * 2 + 1) * Fp + f] : make_float2(0.f, 0.f);
    }

    for (int s0 = 0; s0 < NS; s0 += 2) {  // two sources per complex FFT
        const bool has1 = s0 + 1 < NS;
        float2 X[EPL];
        const size_t r0 = ((size_t)s * NS + s0) * B + l, r1 = r0 + B;
#pragma unroll
        for (int u = 0; u < HB; ++u) {
            X[u].x = prm.xprev[r0 + G * u];
            X[u + HB].x = prm.x[r0 + G * u];
            X[u].y = has1 ? prm.xprev[r1 + G * u] : 0.f;
            X[u + HB].y = has1 ? prm.x[r1 + G * u] : 0.f;
        }
        if (grp == 0) {
#pragma unroll
            for (int u = 0; u < HB; ++u) {
                prm.xnext[r0 + G * u] = X[u + HB].x;
                if (has1) prm.xnext[r1 + G * u] = X[u + HB].y;
            }
        }
        fft_sync<N2>();
        warp_fft<N2, false>(X, buf, tw, l);
        float2 X0[HB + 1], X1[HB + 1];
        float2 M[HB + 1];
        fft_mirror<N2>(X, M, buf, l);
#pragma unroll
        for (int u = 0; u <= HB; ++u) {
            const float2 zf = X[u], zp = M[u];
            X0[u] = make_float2(0.5f * (zf.x + zp.x), 0.5f * (zf.y - zp.y));
            X1[u] = make_float2(0.5f * (zf.y + zp.y), 0.5f * (zp.x - zf.x));
        }
        if (prm.passthrough) {  // queue slot 0 := spectra of the first two sources (_convolver.py:558-562)
            if (s0 == 0) {
#pragma unroll
                for (int u = 0; u <= HB; ++u) { acc[u][0] = X0[u]; if (has1) acc[u][1] = X1[u]; }
            }
            continue;
        }
        // directions of the two sources for the current and the previous block
        const int dc0 = direction_index(yaw, prm.tracker_dir, prm.src_azim16[s0], prm.n_dirs);
        const int dc1 = has1 ? direction_index(yaw, prm.tracker_dir, prm.src_azim16[s0 + 1], prm.n_dirs) : 0;
        const int dl0 = prm.dir_last_in[(size_t)s * NS + s0];
        const int dl1 = has1 ? prm.dir_last_in[(size_t)s * NS + s0 + 1] : 0;
        if (prm.crossfade && (BIG ? threadIdx.x == 0 : lane == 0)) {
            prm.dir_last_out[(size_t)s * NS + s0] = dc0;
            if (has1) prm.dir_last_out[(size_t)s * NS + s0 + 1] = dc1;
        }
        if (HEAD_ONLY && (BIG ? threadIdx.x == 0 : lane == 0)) {
            prm.dir_cur[(size_t)s * NS + s0] = dc0;
            if (has1) prm.dir_cur[(size_t)s * NS + s0 + 1] = dc1;
        }
        const int Pend = HEAD_ONLY ? 1 : prm.P;
        const bool use_last = do_last && prm.last_valid;
        const size_t dstride = (size_t)prm.P * 2 * Fp;
        // Bin slot outermost, partitions in batches of PB: the filter spectra (4 directions x 2 ears)
        // and the queue values of a whole batch are requested before the first one is used, so a lane
        // has 12 * PB loads in flight (the first version walked partition by partition and waited for
        // every load where it was issued: ncu 44 long-scoreboard stall cycles per issue, issue-active
        // 0.04, 0.71 ms for 4 096 streams x 2 sources x 8 partitions).
        const float2 *hd[4] = {prm.hfd + dc0 * dstride, prm.hfd + dc1 * dstride,
                               prm.hfd + dl0 * dstride, prm.hfd + dl1 * dstride};
#pragma unroll
        for (int u = 0; u <= HB; ++u) {
            const int f = (u < HB) ? (l + G * u) : B;
            const bool own = (u < HB) || (l == 0);
            const float2 a0 = make_float2(X0[u].x * prm.inv_ns, X0[u].y * prm.inv_ns);
            const float2 a1 = make_float2(X1[u].x * prm.inv_ns, X1[u].y * prm.inv_ns);
            if (HEAD_ONLY && own && grp == 0) {
                float2 *sp = prm.spec + ((size_t)s * NS + s0) * Fp + f;
                sp[0] = a0;
                if (has1) sp[Fp] = a1;
            }
            for (int p0 = 0; p0 < Pend; p0 += PB) {
                float2 h[PB][4][2], q[PB][4];
                float2 *qp[PB][2];
#pragma unroll
                for (int i = 0; i < PB; ++i) {
                    const int p = min(p0 + i, Pend - 1);  // clamped: surplus batch entries are not used
                    qp[i][0] = qc + (size_t)((prm.head_cur + p) % prm.P) * 2 * Fp + f;
                    qp[i][1] = ql + (size_t)((prm.head_last + p) % prm.P) * 2 * Fp + f;
#pragma unroll
                    for (int d = 0; d < 4; ++d)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const bool need = (d < 2 || use_last) && ((d & 1) == 0 || has1);
                            h[i][d][e] = need ? __ldg(hd[d] + (size_t)p * 2 * Fp + (size_t)e * Fp + f)
                                              : make_float2(0.f, 0.f);
                        }
                    const bool rmw = own && grp == 0 && p0 + i > 0 && p0 + i < Pend;
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        q[i][e] = rmw ? qp[i][0][(size_t)e * Fp] : make_float2(0.f, 0.f);
                        q[i][2 + e] = (rmw && use_last) ? qp[i][1][(size_t)e * Fp] : make_float2(0.f, 0.f);
                    }
                }
#pragma unroll
                for (int i = 0; i < PB; ++i) {
                    const int p = p0 + i;
                    if (p >= Pend) break;
                    float2 c[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        cfma(c[e], h[i][0][e], a0);
                        if (has1) cfma(c[e], h[i][1][e], a1);
                        if (use_last) {
                            cfma(c[2 + e], h[i][2][e], a0);
                            if (has1) cfma(c[2 + e], h[i][3][e], a1);
                        }
                    }
                    if (p == 0) {  // the head stays in registers
#pragma unroll
                        for (int k = 0; k < 4; ++k) { acc[u][k].x += c[k].x; acc[u][k].y += c[k].y; }
                    } else if (own && grp == 0) {
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            qp[i][0][(size_t)e * Fp] = make_float2(q[i][e].x + c[e].x, q[i][e].y + c[e].y);
                            if (use_last)
                                qp[i][1][(size_t)e * Fp] = make_float2(q[i][2 + e].x + c[2 + e].x, q[i][2 + e].y + c[2 + e].y);
                        }
                    }
                }
            }
        }
    }
    // the heads leave the queues
    fft_sync<N2>();
    // (split launch: the consumed slots stay as they are; the next block's queue kernel overwrites them)
    if (!HEAD_ONLY && grp == 0) {
#pragma unroll
        for (int u = 0; u <= HB; ++u) {
            const int f = (u < HB) ? (l + G * u) : B;
            if (u < HB || l == 0) {
                qc[((size_t)prm.head_cur * 2 + 0) * Fp + f] = make_float2(0.f, 0.f);
                qc[((size_t)prm.head_cur * 2 + 1) * Fp + f] = make_float2(0.f, 0.f);
                if (do_last) {
                    ql[((size_t)prm.head_last * 2 + 0) * Fp + f] = make_float2(0.f, 0.f);
                    ql[((size_t)prm.head_last * 2 + 1) * Fp + f] = make_float2(0.f, 0.f);
                }
            }
        }
    }
    render_tail<N2>(acc, buf, tw, grp, l, prm.win_in, prm.win_out, do_last ? 1 : 0,
                    prm.out + (size_t)s * 2 * B);
}

// Split launch, second kernel: the partitions 1 .. P-1 of both output-side queues,
//   q[(head + p) % P][ear][f] += sum_src H[dir_src][p][ear][f] * A_src[f],
// as a streaming read-modify-write: one thread owns two adjacent bins (one float4) of one stream and
// walks the partitions with the source spectra in registers; the filter spectra come from L2 (the
// whole table is 360 x P x 2 x Fp x 8 B).  Every queue byte is read and written once per block
// regardless of the number of sources (the fused kernel passes over the queues once per source pair).
// The slot of partition P-1 is the head the previous block consumed: it is overwritten, not
// accumulated into, so nobody has to clear a consumed head (saves one write and one read of a slot
// per queue and block; the plan clears the slot when a block of another launch variant follows).
// NST = number of sources kept in registers (0: any number, spectra re-read per partition).
#ifndef RTB_FD_HEAD_MINB
#define RTB_FD_HEAD_MINB 3   // resident CTAs per SM of the head kernel of the split launch (164 registers)
#endif
#ifndef RTB_FDQ_QB
#define RTB_FDQ_QB 2
#endif
#ifndef RTB_FDQ_MINB
#define RTB_FDQ_MINB 1
#endif
template <int NST>
__global__ void __launch_bounds__(128, RTB_FDQ_MINB)
fd_queue_kernel(const FdParams prm) {
    const int half = prm.Fp / 2;  // float4 per spectrum row
    const long long item = (long long)blockIdx.x * 128 + threadIdx.x;
    if (item >= (long long)prm.S * half) return;
    const int s = (int)(item / half), i = (int)(item % half);
    const int NS = prm.NS, Fp = prm.Fp, P = prm.P;
    const bool do_last = prm.crossfade != 0, use_last = do_last && prm.last_valid;
    const size_t dstride = (size_t)P * 2 * Fp;
    constexpr int NR = NST > 0 ? NST : 1;
    float4 a[NR];
    const float4 *hc[NR], *hl[NR];
    const float4 *spec = reinterpret_cast<const float4 *>(prm.spec + (size_t)s * NS * Fp) + i;
    if (NST > 0) {
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            a[k] = spec[(size_t)k * half];
            hc[k] = reinterpret_cast<const float4 *>(prm.hfd + prm.dir_cur[(size_t)s * NS + k] * dstride) + i;
            hl[k] = reinterpret_cast<const float4 *>(prm.hfd + prm.dir_last_in[(size_t)s * NS + k] * dstride) + i;
        }
    }
    float4 *qc = reinterpret_cast<float4 *>(prm.qcur + (size_t)s * P * 2 * Fp) + i;
    float4 *ql = reinterpret_cast<float4 *>(prm.qlast + (size_t)s * P * 2 * Fp) + i;
    auto mac2 = [](float4 &c, const float4 h, const float4 x) {  // two complex products
        c.x = fmaf(h.x, x.x, c.x); c.x = fmaf(-h.y, x.y, c.x);
        c.y = fmaf(h.x, x.y, c.y); c.y = fmaf(h.y, x.x, c.y);
        c.z = fmaf(h.z, x.z, c.z); c.z = fmaf(-h.w, x.w, c.z);
        c.w = fmaf(h.z, x.w, c.w); c.w = fmaf(h.w, x.z, c.w);
    };
    // (partition, ear) rows, QB at a time so that QB * (2 + 2 * NST) 16-byte loads are in flight
    constexpr int QB = RTB_FDQ_QB;
    const int rows = (P - 1) * 2;
    for (int r0 = 0; r0 < rows; r0 += QB) {
        float4 vq[QB][2], c[QB][2];
        float4 *pq[QB][2];
#pragma unroll
        for (int j = 0; j < QB; ++j) {
            const int r = min(r0 + j, rows - 1), p = 1 + (r >> 1), e = r & 1;
            pq[j][0] = qc + ((size_t)((prm.head_cur + p) % P) * 2 + e) * half;
            pq[j][1] = ql + ((size_t)((prm.head_last + p) % P) * 2 + e) * half;
            const bool fresh = p == P - 1;
            vq[j][0] = fresh ? make_float4(0.f, 0.f, 0.f, 0.f) : *pq[j][0];
            vq[j][1] = (use_last && !fresh) ? *pq[j][1] : make_float4(0.f, 0.f, 0.f, 0.f);
            c[j][0] = c[j][1] = make_float4(0.f, 0.f, 0.f, 0.f);
            const size_t ho = ((size_t)p * 2 + e) * half;
            if (NST > 0) {
#pragma unroll
                for (int k = 0; k < NR; ++k) {
                    mac2(c[j][0], __ldg(hc[k] + ho), a[k]);
                    if (use_last) mac2(c[j][1], __ldg(hl[k] + ho), a[k]);
                }
            } else {
                for (int k = 0; k < NS; ++k) {
                    const float4 ak = spec[(size_t)k * half];
                    const float4 *hck = reinterpret_cast<const float4 *>(prm.hfd + prm.dir_cur[(size_t)s * NS + k] * dstride) + i;
                    mac2(c[j][0], __ldg(hck + ho), ak);
                    if (use_last) {
                        const float4 *hlk = reinterpret_cast<const float4 *>(prm.hfd + prm.dir_last_in[(size_t)s * NS + k] * dstride) + i;
                        mac2(c[j][1], __ldg(hlk + ho), ak);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < QB; ++j) {
            if (r0 + j >= rows) break;
            *pq[j][0] = make_float4(vq[j][0].x + c[j][0].x, vq[j][0].y + c[j][0].y, vq[j][0].z + c[j][0].z, vq[j][0].w + c[j][0].w);
            if (use_last || (do_last && r0 + j >= rows - 2))  // rows - 2, rows - 1: partition P-1
                *pq[j][1] = make_float4(vq[j][1].x + c[j][1].x, vq[j][1].y + c[j][1].y, vq[j][1].z + c[j][1].z, vq[j][1].w + c[j][1].w);
        }
    }
}

}  // namespace

struct rtb_fd_plan {
    int device = 0;
    int S = 0, B = 0, NS = 0, n_dirs = 0, P = 0, F = 0, Fp = 0, N2 = 0;
    bool filter_set = false;
    float *d_xstate = nullptr;            // [2][S][NS][B]
    float2 *d_hfd = nullptr;              // [n_dirs][P][2][Fp]
    float2 *d_qcur = nullptr, *d_qlast = nullptr;
    int *d_dirlast = nullptr;             // [2][S][NS]
    float2 *d_spec = nullptr;             // [S][NS][Fp] (split launch)
    int *d_dircur = nullptr;              // [S][NS]     (split launch)
    float *d_src = nullptr;               // [NS]
    float2 *d_tw = nullptr;
    float *d_win_in = nullptr, *d_win_out = nullptr;
    float *d_in = nullptr, *d_yaw = nullptr, *d_out = nullptr;
    cudaStream_t own_stream = nullptr;
    int par = 0, dir_idx = 0, head_cur = 0, head_last = 0;
    bool crossfade = true, passthrough = false, last_valid = false;
    float tracker_dir = 1.f;
    long long launches = 0;
    int n_sm = 148, dense_forced = -1;    // RTB_FD_DENSE=0/1 forces one kernel shape (A/B runs)
    int split_forced = -1;                // RTB_FD_SPLIT=0/1 forces the fused / the two-kernel launch
    int stale_cur = -1, stale_last = -1;  // consumed head slots a split block left uncleared
};

namespace {
size_t fd_queue_elems(const rtb_fd_plan *p) { return (size_t)p->S * p->P * 2 * p->Fp; }
size_t fd_state_elems(const rtb_fd_plan *p) { return (size_t)p->S * p->NS * p->B; }
}  // namespace

extern "C" {

int rtb_fd_plan_create(rtb_fd_plan **out, int device, int n_streams, int block_length, int n_sources,
                       int n_dirs, int n_partitions) {
    if (!out) return fail(RTB_ERR_INVALID, "plan output pointer is NULL");
    *out = nullptr;
    if (n_streams < 1 || n_dirs < 1 || n_partitions < 1) return fail(RTB_ERR_INVALID, "sizes must be >= 1");
    // the direction index is int(-azimuth) on the integer-degree SSR grid (_convolver.py:913-945,
    // _filter_set.py:883-899): 0 or n_dirs - k with k = 1 .. 359.  With fewer directions the
    // reference fails with an IndexError at run time; refuse the plan instead of reading out of bounds.
    if (n_dirs < 360)
        return fail(RTB_ERR_INVALID, "the direction-switching renderer indexes filters by whole degrees of "
                    "azimuth and needs at least 360 directions (got %d)", n_dirs);
    if (n_sources < 1)
        return fail(RTB_ERR_INVALID, "no virtual source positions (according to playback file channel count) given.");
    if (!block_length_supported(block_length))
        return fail(RTB_ERR_UNSUPPORTED, "block length %d not supported (supported: " RTB_BLOCK_LENGTHS ")", block_length);
    int ndev = 0;
    RTB_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(RTB_ERR_INVALID, "no CUDA device %d", device);
    rtb_fd_plan *p = new (std::nothrow) rtb_fd_plan();
    if (!p) return fail(RTB_ERR_NOMEM, "out of host memory");
    p->device = device; p->S = n_streams; p->B = block_length; p->NS = n_sources; p->n_dirs = n_dirs;
    p->P = n_partitions; p->F = block_length + 1; p->Fp = (p->F + 3) / 4 * 4; p->N2 = 2 * block_length;
    DeviceGuard guard(device);
    cudaDeviceGetAttribute(&p->n_sm, cudaDevAttrMultiProcessorCount, device);
    if (const char *env = getenv("RTB_FD_DENSE")) p->dense_forced = env[0] != '0';
    if (const char *env = getenv("RTB_FD_SPLIT")) p->split_forced = env[0] != '0';
    cudaError_t err = cudaStreamCreateWithFlags(&p->own_stream, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaMalloc(&p->d_xstate, sizeof(float) * 2 * fd_state_elems(p));
    if (err == cudaSuccess) err = cudaMalloc(&p->d_hfd, sizeof(float2) * (size_t)n_dirs * p->P * 2 * p->Fp);
    if (err == cudaSuccess) err = cudaMalloc(&p->d_qcur, sizeof(float2) * fd_queue_elems(p));
    if (err == cudaSuccess) err = cudaMalloc(&p->d_qlast, sizeof(float2) * fd_queue_elems(p));
    if (err == cudaSuccess) err = cudaMalloc(&p->d_dirlast, sizeof(int) * 2 * (size_t)p->S * p->NS);
    if (err == cudaSuccess) err = cudaMalloc(&p->d_spec, sizeof(float2) * (size_t)p->S * p->NS * p->Fp);
    if (err == cudaSuccess) err = cudaMalloc(&p->d_dircur, sizeof(int) * (size_t)p->S * p->NS);
    if (err == cudaSuccess) err = cudaMemset(p->d_spec, 0, sizeof(float2) * (size_t)p->S * p->NS * p->Fp);
    if (err == cudaSuccess) err = cudaMalloc(&p->d_src, sizeof(float) * p->NS);
    if (err == cudaSuccess) err = cudaMalloc(&p->d_tw, sizeof(float2) * p->N2);
    if (err == cudaSuccess) err = cudaMalloc(&p->d_win_in, sizeof(float) * p->B);
    if (err == cudaSuccess) err = cudaMalloc(&p->d_win_out, sizeof(float) * p->B);
    if (err == cudaSuccess) err = cudaMemset(p->d_src, 0, sizeof(float) * p->NS);
    if (err == cudaSuccess) err = cudaMemset(p->d_dirlast, 0, sizeof(int) * 2 * (size_t)p->S * p->NS);
    if (err != cudaSuccess) {
        rtb_fd_plan_destroy(p);
        return fail(err == cudaErrorMemoryAllocation ? RTB_ERR_NOMEM : RTB_ERR_CUDA,
                    "plan allocation failed: %s", cudaGetErrorString(err));
    }
    std::vector<float2> tw(p->N2);
    for (int q = 0; q < p->N2; ++q) {
        const double a = -2.0 * M_PI * q / p->N2;
        tw[q] = make_float2((float)cos(a), (float)sin(a));
    }
    RTB_CUDA(cudaMemcpy(p->d_tw, tw.data(), sizeof(float2) * p->N2, cudaMemcpyHostToDevice));
    *out = p;
    return rtb_fd_plan_reset(p);
}

int rtb_fd_plan_set_filter(rtb_fd_plan *p, const float *hfd, const float *win_in, const float *win_out) {
    if (!p || !hfd) return fail(RTB_ERR_INVALID, "NULL argument");
    DeviceGuard guard(p->device);
    // reference layout [P][dirs][2][F] -> device layout [dirs][P][2][Fp]
    std::vector<float2> h((size_t)p->n_dirs * p->P * 2 * p->Fp, make_float2(0.f, 0.f));
    for (int q = 0; q < p->P; ++q)
        for (int d = 0; d < p->n_dirs; ++d)
            for (int e = 0; e < 2; ++e)
                for (int f = 0; f < p->F; ++f) {
                    const size_t i = ((((size_t)q * p->n_dirs + d) * 2 + e) * p->F + f);
                    h[(((size_t)d * p->P + q) * 2 + e) * p->Fp + f] = make_float2(hfd[2 * i], hfd[2 * i + 1]);
                }
    std::vector<float> wi(p->B), wo(p->B);
    for (int n = 0; n < p->B; ++n) {
        float o, i;
        if (win_in && win_out) { i = win_in[n]; o = win_out[n]; }
        else {
            const double c0 = cos(n * M_PI / (2.0 * (p->B - 1))), c1 = cos((p->B - 1 - n) * M_PI / (2.0 * (p->B - 1)));
            o = (float)(c0 * c0); i = (float)(c1 * c1);
        }
        wi[n] = i / (float)p->N2; wo[n] = o / (float)p->N2;
    }
    RTB_CUDA(cudaDeviceSynchronize());
    RTB_CUDA(cudaMemcpy(p->d_hfd, h.data(), sizeof(float2) * h.size(), cudaMemcpyHostToDevice));
    RTB_CUDA(cudaMemcpy(p->d_win_in, wi.data(), sizeof(float) * p->B, cudaMemcpyHostToDevice));
    RTB_CUDA(cudaMemcpy(p->d_win_out, wo.data(), sizeof(float) * p->B, cudaMemcpyHostToDevice));
    RTB_CUDA(cudaDeviceSynchronize());  // uploads and creation-time memsets done before a non-blocking stream runs
    p->filter_set = true;
    return RTB_OK;
}

int rtb_fd_plan_set_sources(rtb_fd_plan *p, const float *source_azim_deg, int tracker_dir) {
    if (!p || !source_azim_deg) return fail(RTB_ERR_INVALID, "NULL argument");
    if (tracker_dir != 1 && tracker_dir != -1) return fail(RTB_ERR_INVALID, "tracker_dir must be +1 or -1");
    DeviceGuard guard(p->device);
    std::vector<float> src(p->NS);
    for (int i = 0; i < p->NS; ++i) src[i] = __half2float(__float2half_rn(source_azim_deg[i]));
    RTB_CUDA(cudaMemcpy(p->d_src, src.data(), sizeof(float) * p->NS, cudaMemcpyHostToDevice));
    p->tracker_dir = (float)tracker_dir;
    return RTB_OK;
}

int rtb_fd_plan_set_passthrough(rtb_fd_plan *p, int on) {
    if (!p) return fail(RTB_ERR_INVALID, "NULL plan");
    p->passthrough = on != 0;
    return RTB_OK;
}

int rtb_fd_plan_set_crossfade(rtb_fd_plan *p, int on) {
    if (!p) return fail(RTB_ERR_INVALID, "NULL plan");
    p->crossfade = on != 0;
    if (!p->crossfade) {  // _last_blocks_fd.fill(0); _last_filters_fd.fill(0)  (_convolver.py:887-890)
        DeviceGuard guard(p->device);
        RTB_CUDA(cudaDeviceSynchronize());
        RTB_CUDA(cudaMemset(p->d_qlast, 0, sizeof(float2) * fd_queue_elems(p)));
        RTB_CUDA(cudaDeviceSynchronize());
        p->last_valid = false;
        p->stale_last = -1;
    }
    return RTB_OK;
}

int rtb_fd_plan_reset(rtb_fd_plan *p) {
    if (!p) return fail(RTB_ERR_INVALID, "NULL plan");
    DeviceGuard guard(p->device);
    RTB_CUDA(cudaDeviceSynchronize());  // blocks in flight on non-blocking streams (see rtb_ols_plan_reset)
    RTB_CUDA(cudaMemset(p->d_xstate, 0, sizeof(float) * 2 * fd_state_elems(p)));
    RTB_CUDA(cudaMemset(p->d_qcur, 0, sizeof(float2) * fd_queue_elems(p)));
    RTB_CUDA(cudaMemset(p->d_qlast, 0, sizeof(float2) * fd_queue_elems(p)));
    RTB_CUDA(cudaDeviceSynchronize());
    p->stale_cur = p->stale_last = -1;
    return RTB_OK;
}

int rtb_fd_process_block(rtb_fd_plan *p, const float *d_in, const float *d_yaw_deg, float *d_out, void *cuda_stream) {
    if (!p || !d_in || !d_out) return fail(RTB_ERR_INVALID, "NULL argument");
    if (!p->filter_set) return fail(RTB_ERR_STATE, "blocks in frequency domain have not been calculated yet.");
    if (!d_yaw_deg && !p->passthrough) return fail(RTB_ERR_INVALID, "NULL yaw pointer");
    DeviceGuard guard(p->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    FdParams prm;
    prm.x = d_in;
    prm.xprev = p->d_xstate + (size_t)p->par * fd_state_elems(p);
    prm.xnext = p->d_xstate + (size_t)(p->par ^ 1) * fd_state_elems(p);
    prm.hfd = p->d_hfd; prm.qcur = p->d_qcur; prm.qlast = p->d_qlast;
    prm.yaw_deg = d_yaw_deg ? d_yaw_deg : p->d_src;  // unused in passthrough
    prm.src_azim16 = p->d_src;
    prm.dir_last_in = p->d_dirlast + (size_t)p->dir_idx * p->S * p->NS;
    prm.dir_last_out = p->d_dirlast + (size_t)(p->dir_idx ^ 1) * p->S * p->NS;
    prm.tw = p->d_tw; prm.win_in = p->d_win_in; prm.win_out = p->d_win_out; prm.out = d_out;
    prm.S = p->S; prm.NS = p->NS; prm.P = p->P; prm.Fp = p->Fp; prm.n_dirs = p->n_dirs;
    prm.head_cur = p->head_cur; prm.head_last = p->head_last;
    prm.crossfade = p->crossfade ? 1 : 0; prm.last_valid = p->last_valid ? 1 : 0;
    prm.passthrough = p->passthrough ? 1 : 0;
    prm.tracker_dir = p->tracker_dir; prm.inv_ns = 1.0f / (float)p->NS;
    prm.spec = p->d_spec; prm.dir_cur = p->d_dircur;
    // Two shapes of the same kernel (4 096 streams x 2 sources x 8 partitions at B = 256, one B200):
    //   3 partitions per load batch, registers unlimited (243, 2 CTAs/SM):
    //       0.0486 ms @1024, 0.104 @2048, 0.154 @3072, 0.202 @4096, 0.364 @8192 streams
    //   2 partitions per batch, 128 registers (4 CTAs/SM, 200 B stack):
    //       0.0647 ms @1024, 0.094 @2048, 0.159 @3072, 0.182 @4096, 0.351 @8192
    // (4 per batch, the first version: 255 registers + spill, 0.348 ms @4096; 1 per batch 0.285 ms;
    // 8 per batch 0.72 ms.)  The dense shape pays off once most SMs hold their four CTAs.
    const bool dense = p->dense_forced >= 0 ? p->dense_forced != 0
                                            : (p->N2 <= 512 && (long long)(p->S + 3) / 4 * 10 >= 34LL * p->n_sm);
    // Split launch (default whenever there are partitions behind the head): head kernel + streaming
    // queue kernel.  Same workload as above, fused -> split: 0.0329 -> 0.0185 ms @256 streams,
    // 0.0492 -> 0.0376 @1024, 0.182 -> 0.124 @4096, 0.350 -> 0.227 @8192; 32 partitions @4096
    // 0.894 -> 0.456 ms.  (Queue kernel 2 rows per batch / 94 registers; 4 rows 0.147 vs 0.137 ms,
    // 8 rows 0.159; head kernel 3 CTAs per SM 0.124 vs 0.127 with 2.)
    const bool split = !p->passthrough && p->P > 1 &&
                       (p->split_forced >= 0 ? p->split_forced != 0 : true);
    if (!split) {  // the fused kernel accumulates into every slot: clear what a split block left behind
        const size_t pitch = sizeof(float2) * (size_t)p->P * 2 * p->Fp, width = sizeof(float2) * 2 * p->Fp;
        if (p->stale_cur >= 0)
            RTB_CUDA(cudaMemset2DAsync(p->d_qcur + (size_t)p->stale_cur * 2 * p->Fp, pitch, 0, width, p->S, st));
        if (p->stale_last >= 0)
            RTB_CUDA(cudaMemset2DAsync(p->d_qlast + (size_t)p->stale_last * 2 * p->Fp, pitch, 0, width, p->S, st));
        p->stale_cur = p->stale_last = -1;
    } else {
        p->stale_cur = p->head_cur;
        if (p->crossfade) p->stale_last = p->head_last;
    }
#define RTB_FD_LAUNCH(N)                                                                             \
    if (split) fft_kernel_launch<N>(fd_render_kernel<N, 1, FftCfg<N>::BIG ? 1 : RTB_FD_HEAD_MINB, true>, prm, p->S, st, 0, 4); \
    else if (FftCfg<N>::BIG || !dense) fft_kernel_launch<N>(fd_render_kernel<N, 3, 1>, prm, p->S, st, 0, 4); \
    else fft_kernel_launch<N>(fd_render_kernel<N, 2, FftCfg<N>::BIG ? 1 : 4>, prm, p->S, st, 0, 4);
    RTB_DISPATCH_N2(p->N2, RTB_FD_LAUNCH)
#undef RTB_FD_LAUNCH
    if (split) {
        const long long items = (long long)p->S * (p->Fp / 2);
        const unsigned grid = (unsigned)((items + 127) / 128);
        switch (p->NS) {
            case 1: fd_queue_kernel<1><<<grid, 128, 0, st>>>(prm); break;
            case 2: fd_queue_kernel<2><<<grid, 128, 0, st>>>(prm); break;
            case 3: fd_queue_kernel<3><<<grid, 128, 0, st>>>(prm); break;
            case 4: fd_queue_kernel<4><<<grid, 128, 0, st>>>(prm); break;
            default: fd_queue_kernel<0><<<grid, 128, 0, st>>>(prm); break;
        }
        p->launches++;
    }
    p->launches++;
    p->par ^= 1;
    p->head_cur = (p->head_cur + 1) % p->P;
    if (!p->passthrough && p->crossfade) {  // _last_filters_fd = filters (_convolver.py:829-831)
        p->head_last = (p->head_last + 1) % p->P;
        p->dir_idx ^= 1;
        p->last_valid = true;
    }
    RTB_CUDA(cudaGetLastError());
    return RTB_OK;
}

int rtb_fd_process_block_host(rtb_fd_plan *p, const float *h_in, const float *h_yaw_deg, float *h_out) {
    if (!p || !h_in || !h_out) return fail(RTB_ERR_INVALID, "NULL argument");
    DeviceGuard guard(p->device);
    const size_t n_in = fd_state_elems(p), n_out = (size_t)p->S * 2 * p->B;
    if (!p->d_in) {
        cudaError_t err = cudaMalloc(&p->d_in, sizeof(float) * n_in);
        if (err == cudaSuccess) err = cudaMalloc(&p->d_out, sizeof(float) * n_out);
        if (err == cudaSuccess) err = cudaMalloc(&p->d_yaw, sizeof(float) * p->S);
        if (err == cudaSuccess) err = cudaMemset(p->d_yaw, 0, sizeof(float) * p->S);
        if (err != cudaSuccess) return fail(RTB_ERR_NOMEM, "staging allocation failed: %s", cudaGetErrorString(err));
    }
    RTB_CUDA(cudaMemcpyAsync(p->d_in, h_in, sizeof(float) * n_in, cudaMemcpyHostToDevice, p->own_stream));
    if (h_yaw_deg)
        RTB_CUDA(cudaMemcpyAsync(p->d_yaw, h_yaw_deg, sizeof(float) * p->S, cudaMemcpyHostToDevice, p->own_stream));
    int rc = rtb_fd_process_block(p, p->d_in, p->d_yaw, p->d_out, p->own_stream);
    if (rc != RTB_OK) return rc;
    RTB_CUDA(cudaMemcpyAsync(h_out, p->d_out, sizeof(float) * n_out, cudaMemcpyDeviceToHost, p->own_stream));
    RTB_CUDA(cudaStreamSynchronize(p->own_stream));
    return RTB_OK;
}

int rtb_fd_plan_query(rtb_fd_plan *p, int what, int64_t *value) {
    if (!p || !value) return fail(RTB_ERR_INVALID, "NULL argument");
    switch (what) {
        case RTB_Q_LAUNCHES: *value = p->launches; break;
        case RTB_Q_STATE_BYTES:
            *value = (int64_t)sizeof(float2) * 2 * p->P * 2 * p->Fp + (int64_t)sizeof(float) * 2 * p->NS * p->B;
            break;
        default: return fail(RTB_ERR_INVALID, "unknown query %d", what);
    }
    return RTB_OK;
}

int rtb_fd_plan_destroy(rtb_fd_plan *p) {
    if (!p) return RTB_OK;
    DeviceGuard guard(p->device);
    cudaFree(p->d_xstate); cudaFree(p->d_hfd); cudaFree(p->d_qcur); cudaFree(p->d_qlast);
    cudaFree(p->d_dirlast); cudaFree(p->d_src); cudaFree(p->d_tw); cudaFree(p->d_win_in); cudaFree(p->d_win_out);
    cudaFree(p->d_in); cudaFree(p->d_yaw); cudaFree(p->d_out); cudaFree(p->d_spec); cudaFree(p->d_dircur);
    if (p->own_stream) cudaStreamDestroy(p->own_stream);
    delete p;
    return RTB_OK;
}

}  // extern "C"
