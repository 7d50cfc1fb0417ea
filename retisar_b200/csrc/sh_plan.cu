// Host side of the SH-domain renderer: table preparation, per-stream state, launches, C ABI.
// Reference behaviour: ReTiSAR/_convolver.py AdjustableShConvolver (:978-1341) and the parts of
// AdjustableFdConvolver / OverlapSaveConvolver it inherits.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <new>
#include <vector>

#include "sh_kernels.cuh"

using namespace rtb;

#define RTB_MAX_CHUNKS 4
#define RTB_HOST_SLOTS 3

struct rtb_sh_plan {
    int device = 0;
    int S = 0, B = 0, C = 0, order = 0, K = 0, E = 2, P = 1;
    int F = 0, N2 = 0;
    // derived by set_tables
    bool tables_set = false;
    bool symmetric = false;
    int J = 0, Jpad = 0, JT = 0, TT = 0, n_jchunks = 0, NP = 0;
    // bulk-copy analysis kernel: tile width, pipeline depth, grid, dynamic smem (0 = not usable)
    int an_TW = 0, an_stages = 0, an_grid = 0, an_smem = 0, an_nrun = 2;
    bool an_rsmem = true;
    // tensor-core analysis kernel (wide arrays): padded sizes, operand table, launch config
    bool tc_enabled = false, tc_stream_b = false;
    // warp-level tensor-core analysis kernel (narrow arrays, sh_analysis_mma.cuh)
    bool mma_enabled = false;
    int mma_TW = 0, mma_jchunks = 0, mma_smem = 0, mma_grid = 0, mma_stages = 2;
    float *d_mma_r = nullptr;     // [n_jchunks][hi, lo][32 rows][32 channels]
    int tc_Kpad = 0, tc_Npad = 0, tc_KC = 0, tc_smem = 0, tc_cols = 0, tc_grid = 0, tc_nst = 2;
    float *d_tc_b = nullptr;      // [2][Kpad/4][Npad/8][8][4] analysis matrix split into TF32 hi / lo
    // device memory
    float *d_rmT = nullptr;       // [C][Jpad]
    float *d_zring = nullptr;     // [2][S][J][B]
    float4 *d_htab = nullptr;     // [NP][2][F]
    ShSignal *d_signals = nullptr;
    // steady-state render kernel (signals grouped by m, sh_render_mg.cuh): tables for `mg_order`
    // programmatic dependent launch of the chain's kernels: RTB_PDL=1 enables.  Measured on B200:
    // 1 024 streams 55.4 -> 62.8 us per block (worse), 10 240 streams 385 -> 380 us: off by default
    bool pdl = false;
    bool mg_enabled = false, mg_forced = false;
    int n_sm = 148;
    int mg_order = -1, mg_np = 0;
    float4 *d_mg_htab = nullptr;
    ShMgSignal *d_mg_signals = nullptr;
    std::vector<float> h_rows;         // analysis rows [J][C] (rtb_sh_plan_get_analysis_rows)
    std::vector<ShSignal> h_sig;       // host copy of the signal list
    std::vector<float> h_hnm;          // host copy of H_nm [K][2][F] complex (P == 1)
    std::vector<int16_t> h_shm;        // sh_m[K]
    ShStep *d_steps = nullptr;    // flattened (signal, term) list for the partitioned kernels
    std::vector<int> steps_prefix;  // steps_prefix[np] = number of steps of the first np signals
    float2 *d_tw = nullptr;       // [N2]
    float *d_win_in = nullptr, *d_win_out = nullptr;
    float *d_radlast = nullptr;   // [2][S]
    // partitioned HRIR (P > 1): spectra scratch and the two output-side queues
    int Fp = 0;
    float2 *d_zspec = nullptr;    // [S][NP][N2]
    float2 *d_qcur = nullptr, *d_qlast = nullptr;  // [S][P][2][Fp]
    // tensor-core partition MAC (sh_partitioned_tc.cuh): stream-minor layout of the spectra
    // [N2][NP][Spad] and of the queues [F][P][2][Spad], split H operand per bin and K-chunk
    bool pt_enabled = false;
    long long pt_Spad = 0;
    int pt_Npad = 0, pt_acc_cols = 0, pt_chunks_max = 0, pt_smem = 0, pt_grid = 0;
    float *d_pt_b = nullptr;
    const float *pt_hnm_src = nullptr;
    int head_cur = 0, head_last = 0;
    int pt_stale_cur = -1, pt_stale_last = -1;  // consumed heads the tcgen05 partition path left uncleared
    bool pt_overwrite = true;                   // RTB_PT_OVERWRITE=0: clear consumed heads in the tail kernel instead
    int spec_split_forced = 0;                  // RTB_SPEC_SPLIT=n: grid y of the spectra kernel (A/B runs)
    // CUDA graphs of runs of blocks (rtb_sh_process_blocks): key of a run -> instantiated graph
    struct BlockGraph {
        const float *d_in, *d_yaw; float *d_out; int n_ring, phase, count; cudaStream_t st;
        int slot, rad_idx, last_valid, cur_order, last_order, crossfade, passthrough, hp_slot, hp_xidx, mg_order;
        long long launches;          // kernel launches of one replay
        int seen;                    // calls with this key so far (the graph is built on the second)
        cudaGraphExec_t exec;
    };
    std::vector<BlockGraph> graphs;
    bool graphs_enabled = true;                 // RTB_GRAPH=0: every block of a run is launched by the host
    long long graph_replays = 0;
    // staging for the host-buffer entry points: RTB_HOST_SLOTS device copies of (input block, yaw,
    // output block) so that the input copy of block t+1 overlaps the kernels and the result copy
    // of block t (rtb_sh_submit_block_host / rtb_sh_wait_block_host)
    struct HostSlot {
        float *d_in = nullptr, *d_yaw = nullptr, *d_out = nullptr;
        cudaEvent_t in_done = nullptr, k_done = nullptr, out_done = nullptr;
    } hslot[RTB_HOST_SLOTS];
    bool hslots_ready = false;
    long long submitted = 0;      // blocks submitted through the host entry points (ticket counter)
    cudaStream_t own_stream = nullptr, d2h_stream = nullptr, k_stream = nullptr;
    // optional headphone-compensation filter behind the renderer (hpcf_stage.cuh): partitions,
    // filter spectra [P][2][Fp], delay line of binaural spectra [S][P][2][Fp] (ring: hp_slot),
    // binaural block state [2][S][2][B] (ping-pong: hp_xidx receives the current block)
    int hp_P = 0, hp_slot = 0, hp_xidx = 0;
    float2 *d_hp_hfd = nullptr, *d_hp_hist = nullptr;
    float *d_hp_x = nullptr;
    // control state (reference attribute it mirrors)
    int slot = 0;                 // which half of d_zring receives the current block
    int rad_idx = 0;              // which half of d_radlast holds the previous yaw
    int cur_order = 0;            // _sh_cur_order
    int last_order = 0;           // length of _last_sh_azim_nm, as an order
    bool last_valid = false;      // false <=> _last_sh_azim_nm is all zeros
    bool crossfade = true;        // _is_crossfade
    bool passthrough = false;     // _is_passthrough
    std::vector<int> sig_k1, sig_k2;   // H_nm row consumed by term 1 / term 2 of each signal
    std::vector<float> sig_sign2;      // (-1)^m folded into term 2
    float src_azim16 = 0.f;       // _sources_deg[0, 0], already rounded to float16
    float tracker_dir = 1.f;      // +1 HRIR, -1 BRIR
    long long launches = 0;
    int last_render_kernel = 0;   // RTB_Q_RENDER_KERNEL
    // optional per-kernel timing (CUDA events on the launching stream)
    bool profiling = false;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    bool ev_valid = false;
};


static void drop_graphs(rtb_sh_plan *p) {  // recorded runs of blocks hold kernel parameters (table pointers,
    for (auto &e : p->graphs)                // source azimuth, ...): forget them when those change
        if (e.exec) cudaGraphExecDestroy(e.exec);
    p->graphs.clear();
}

namespace {

float round_to_half(float v) { return __half2float(__float2half_rn(v)); }

int signals_for_order(const rtb_sh_plan *p, int order) {
    return p->symmetric ? (order + 1) * (order + 2) / 2 : (order + 1) * (order + 1);
}

template <int JTH, int NRUN, bool R_SMEM>
void launch_analysis_t(const rtb_sh_plan *p, const float *d_in, float *zcur, int count, cudaStream_t st) {
    kernel_attrs_once(sh_analysis_kernel<JTH, NRUN, R_SMEM>, 220 * 1024);
    const long long tiles = (long long)count * (p->B / p->an_TW);
    const int grid = (int)std::min<long long>(tiles, (long long)p->an_grid);
    launch_pdl(sh_analysis_kernel<JTH, NRUN, R_SMEM>, dim3(grid), dim3(128), (size_t)p->an_smem, st, p->pdl,
               d_in, (const float *)p->d_rmT, zcur, count, p->C, p->J, p->B, p->an_TW, p->n_jchunks, p->an_stages);
}

template <int JTH>
void launch_analysis_j(const rtb_sh_plan *p, const float *d_in, float *zcur, int count, cudaStream_t st) {
    if (p->an_nrun == 1) {
        if (p->an_rsmem) launch_analysis_t<JTH, 1, true>(p, d_in, zcur, count, st);
        else launch_analysis_t<JTH, 1, false>(p, d_in, zcur, count, st);
    } else {
        if (p->an_rsmem) launch_analysis_t<JTH, 2, true>(p, d_in, zcur, count, st);
        else launch_analysis_t<JTH, 2, false>(p, d_in, zcur, count, st);
    }
}

void launch_analysis(const rtb_sh_plan *p, const float *d_in, float *zcur, int count, cudaStream_t st) {
    if (p->mma_enabled) {
        kernel_attrs_once(sh_analysis_mma_kernel, 200 * 1024);
        const long long tiles = (long long)count * (p->B / p->mma_TW);
        const int grid = (int)std::min<long long>(tiles, (long long)p->mma_grid);
        sh_analysis_mma_kernel<<<grid, 128, p->mma_smem, st>>>(d_in, p->d_mma_r, zcur, count, p->C, p->J, p->B,
                                                              p->mma_TW, p->mma_jchunks, p->mma_stages);
        return;
    }
    if (p->tc_enabled) {
        const long long tiles = ((long long)count * p->B + RTB_TC_M - 1) / RTB_TC_M;
        const int grid = (int)std::min<long long>(tiles, (long long)p->tc_grid);
        // six loader groups when every CTA has many tiles to stream (measured: c5 8 192 streams 343 vs
        // 356 us, c2 10 240 streams 99 vs 106 us; at 1 024 streams the four-group kernel is faster)
        const bool big = !p->tc_stream_b && tiles >= 32LL * grid;
#define RTB_TC_LAUNCH_T(SB, LGV)                                                                     \
    {                                                                                                \
        kernel_attrs_once(sh_analysis_tc_kernel<SB, LGV>, 224 * 1024);                               \
        sh_analysis_tc_kernel<SB, LGV><<<grid, RTB_TC_THREADS_OF(LGV), p->tc_smem, st>>>(            \
            d_in, p->d_tc_b, zcur, count, p->C, p->J, p->B, p->tc_Kpad, p->tc_Npad, p->tc_cols, p->tc_nst, 0LL); \
    }
        if (p->tc_stream_b) RTB_TC_LAUNCH_T(true, RTB_TC_LG)
        else if (big) RTB_TC_LAUNCH_T(false, RTB_TC_LG_BIG)
        else RTB_TC_LAUNCH_T(false, RTB_TC_LG)
#undef RTB_TC_LAUNCH_T
        return;
    }
    switch (p->JT) {
        case 4: launch_analysis_j<4>(p, d_in, zcur, count, st); break;
        case 5: launch_analysis_j<5>(p, d_in, zcur, count, st); break;
        case 6: launch_analysis_j<6>(p, d_in, zcur, count, st); break;
        case 7: launch_analysis_j<7>(p, d_in, zcur, count, st); break;
        default: launch_analysis_j<8>(p, d_in, zcur, count, st); break;
    }
}

void launch_render(const rtb_sh_plan *p, const ShRenderParams &prm, cudaStream_t st) {
#define RTB_SH_LAUNCH(N) fft_kernel_launch<N>(sh_render_kernel<N>, prm, prm.S, st, 0, 4)
    RTB_DISPATCH_N2(p->N2, RTB_SH_LAUNCH)
#undef RTB_SH_LAUNCH
}

template <int N2, int NW = 4>
void launch_render_mg_t(const ShRenderMgParams &prm, cudaStream_t st, bool pdl) {
    const size_t smem = RTB_MG_ACC_SMEM ? sizeof(float2) * NW * (FftCfg<N2>::EPL / 2) * 4 * 32 : 0;
    constexpr int per_cta = NW * FftCfg<N2>::GROUPS;  // one stream per transform group
    kernel_attrs_once(sh_render_mg_kernel<N2, NW>, 100 * 1024, true);
    launch_pdl(sh_render_mg_kernel<N2, NW>, dim3((unsigned)((prm.S + per_cta - 1) / per_cta)), dim3(32 * NW), smem, st, pdl, prm);
}

void launch_render_mg(const rtb_sh_plan *p, const ShRenderMgParams &prm, cudaStream_t st) {
    // short blocks put 2 or 4 streams on a warp: CTAs of two warps while 4-warp CTAs would not give every
    // SM two of them
    const bool narrow = prm.S < 2 * p->n_sm * 4 * (512 / p->N2 > 2 ? 4 : 2);
    switch (p->N2) {
        case 64: narrow ? launch_render_mg_t<64, 2>(prm, st, p->pdl) : launch_render_mg_t<64>(prm, st, p->pdl); break;
        case 128: narrow ? launch_render_mg_t<128, 2>(prm, st, p->pdl) : launch_render_mg_t<128>(prm, st, p->pdl); break;
        case 256: launch_render_mg_t<256>(prm, st, p->pdl); break;
        default: launch_render_mg_t<512>(prm, st, p->pdl); break;
    }
}

// Tables of the steady-state render kernel for SH order `order`: signals sorted by m (one group
// per phase pair), real m = 0 rows paired two per FFT, term-2 spectra stored mirrored.
int build_mg_tables(rtb_sh_plan *p, int order) {
    drop_graphs(p);  // recorded runs hold the pointers of the tables that are replaced here
    const int F = p->F, B = p->B, K = p->K;
    auto H = [&](int k, int e, int f) {
        const size_t o = 2 * (((size_t)k * 2 + e) * F + f);
        return make_float2(p->h_hnm[o], p->h_hnm[o + 1]);
    };
    std::vector<ShMgSignal> sig;
    std::vector<float4> htab;
    auto add = [&](int row_re, int row_im, int m1, int m2, int has2) {
        ShMgSignal g;
        g.row_re = (short)row_re; g.row_im = (short)row_im;
        g.m1 = (signed char)m1; g.m2 = (signed char)(m2 == RTB_NO_TERM ? RTB_MG_NO_TERM : m2);
        g.has2 = (unsigned char)has2; g.flush = 0;
        sig.push_back(g);
        htab.resize(htab.size() + 2 * (size_t)F, make_float4(0.f, 0.f, 0.f, 0.f));
        return htab.data() + htab.size() - 2 * (size_t)F;
    };
    if (p->symmetric) {
        auto q_of = [](int n, int m) { return n * (n + 1) / 2 + m; };  // index into h_sig
        // m = 0: real rows, two per transform
        for (int n = 0; n <= order; n += 2) {
            const int ka = n * n + n, kb = (n + 1) * (n + 1) + (n + 1);
            const bool pair = n + 1 <= order;
            float4 *t = add(p->h_sig[q_of(n, 0)].row_re, pair ? p->h_sig[q_of(n + 1, 0)].row_re : -1, 0,
                            pair ? 0 : RTB_NO_TERM, pair ? 1 : 0);
            for (int f = 0; f < F; ++f) {
                float2 a[2], b[2];
                for (int e = 0; e < 2; ++e) { a[e] = H(ka, e, f); b[e] = pair ? H(kb, e, f) : make_float2(0.f, 0.f); }
                if (!pair) { t[f] = make_float4(a[0].x, a[0].y, a[1].x, a[1].y); continue; }
                // H1 = (Ha - i Hb) / 2, H2 = (Ha + i Hb) / 2
                t[f] = make_float4(0.5f * (a[0].x + b[0].y), 0.5f * (a[0].y - b[0].x),
                                   0.5f * (a[1].x + b[1].y), 0.5f * (a[1].y - b[1].x));
                t[F + (B - f)] = make_float4(0.5f * (a[0].x - b[0].y), 0.5f * (a[0].y + b[0].x),
                                             0.5f * (a[1].x - b[1].y), 0.5f * (a[1].y + b[1].x));
            }
        }
        sig.back().flush = 1;
        if (order >= 1) sig.back().m2 = 0;  // the group has a V sum even when its last transform is unpaired
        for (int m = 1; m <= order; ++m) {
            for (int n = m; n <= order; ++n) {
                const int kpos = n * n + n + m, kneg = n * n + n - m;
                const ShSignal &s0 = p->h_sig[q_of(n, m)];
                float4 *t = add(s0.row_re, s0.row_im, -m, m, 1);
                const float sg = (m & 1) ? -1.f : 1.f;
                for (int f = 0; f < F; ++f) {
                    const float2 a0 = H(kneg, 0, f), a1 = H(kneg, 1, f), b0 = H(kpos, 0, f), b1 = H(kpos, 1, f);
                    t[f] = make_float4(a0.x, a0.y, a1.x, a1.y);
                    t[F + (B - f)] = make_float4(sg * b0.x, sg * b0.y, sg * b1.x, sg * b1.y);
                }
            }
            sig.back().flush = 1;
        }
    } else {
        const int Ko = (order + 1) * (order + 1);
        std::vector<int> ks(Ko);
        for (int k = 0; k < Ko; ++k) ks[k] = k;
        std::stable_sort(ks.begin(), ks.end(), [&](int a, int b) { return p->h_shm[a] < p->h_shm[b]; });
        for (int i = 0; i < Ko; ++i) {
            const int k = ks[i];
            float4 *t = add(p->h_sig[k].row_re, p->h_sig[k].row_im, p->h_shm[k], RTB_NO_TERM, 0);
            for (int f = 0; f < F; ++f) {
                const float2 a0 = H(k, 0, f), a1 = H(k, 1, f);
                t[f] = make_float4(a0.x, a0.y, a1.x, a1.y);
            }
            if (i + 1 == Ko || p->h_shm[ks[i + 1]] != p->h_shm[k]) sig.back().flush = 1;
        }
    }
    (void)K;
    cudaFree(p->d_mg_htab); p->d_mg_htab = nullptr;
    cudaFree(p->d_mg_signals); p->d_mg_signals = nullptr;
    p->mg_order = -1;
    RTB_CUDA(cudaMalloc(&p->d_mg_htab, sizeof(float4) * htab.size()));
    RTB_CUDA(cudaMalloc(&p->d_mg_signals, sizeof(ShMgSignal) * sig.size()));
    RTB_CUDA(cudaMemcpy(p->d_mg_htab, htab.data(), sizeof(float4) * htab.size(), cudaMemcpyHostToDevice));
    RTB_CUDA(cudaMemcpy(p->d_mg_signals, sig.data(), sizeof(ShMgSignal) * sig.size(), cudaMemcpyHostToDevice));
    p->mg_np = (int)sig.size();
    p->mg_order = order;
    if (p->mg_np > RTB_MG_MAX_SIGNALS) {  // beyond the kernel's shared-memory signal list
        p->mg_enabled = false;
        p->mg_order = -1;
    }
    return RTB_OK;
}

// H table: [P][NP][2][Fs] x (ear0.re, ear0.im, ear1.re, ear1.im); hnm is [P][K][E=2][F] complex.
// Row stride Fs = F for the single-partition kernel (staged by cp.async), Fp (sector aligned) for
// the partitioned one.
std::vector<float4> build_htab(const rtb_sh_plan *p, const float *hnm) {
    const int F = p->F, NP = (int)p->sig_k1.size(), K = p->K;
    const int Fs = p->P > 1 ? p->Fp : F;
    std::vector<float4> htab((size_t)p->P * NP * 2 * Fs, make_float4(0.f, 0.f, 0.f, 0.f));
    for (int part = 0; part < p->P; ++part) {
        auto HNM = [&](int k, int e, int f, int ri) {
            return hnm[2 * ((((size_t)part * K + k) * 2 + e) * F + f) + ri];
        };
        for (int q = 0; q < NP; ++q) {
            const int k1 = p->sig_k1[q], k2 = p->sig_k2[q];
            const float sg = p->sig_sign2[q];
            float4 *row = htab.data() + (((size_t)part * NP + q) * 2) * Fs;
            for (int f = 0; f < F; ++f) {
                row[f] = make_float4(HNM(k1, 0, f, 0), HNM(k1, 0, f, 1), HNM(k1, 1, f, 0), HNM(k1, 1, f, 1));
                if (k2 >= 0)
                    row[Fs + f] = make_float4(sg * HNM(k2, 0, f, 0), sg * HNM(k2, 0, f, 1),
                                              sg * HNM(k2, 1, f, 0), sg * HNM(k2, 1, f, 1));
            }
        }
    }
    return htab;
}

size_t queue_elems(const rtb_sh_plan *p) {
    if (p->pt_enabled) return (size_t)p->F * p->P * 2 * (size_t)p->pt_Spad;
    return (size_t)p->S * p->P * 2 * p->Fp;
}

// Split H operand of the tensor-core partition MAC: per bin f and K-chunk the matrix
// B[n = ((q*2 + ear)*2 + re/im)][k = 2*step + re/im] of the real embedding of the complex product,
//   k even (a_re): (Hr, Hi);  k odd (a_im): (-Hi, Hr),
// in the K-major core-matrix layout [k/4][n/8][n%8][k%4], TF32 hi block then lo block.
std::vector<float> build_pt_b(const rtb_sh_plan *p, const float *hnm) {
    const int F = p->F, K = p->K, P = p->P, Npad = p->pt_Npad, KC = RTB_PT_KC;
    const int NP = (int)p->sig_k1.size();
    struct Step { int k; float sg; };
    std::vector<Step> steps;
    for (int q = 0; q < NP; ++q) {
        steps.push_back({p->sig_k1[q], 1.f});
        if (p->sig_k2[q] >= 0) steps.push_back({p->sig_k2[q], p->sig_sign2[q]});
    }
    const size_t per_chunk = (size_t)2 * KC * Npad, per_bin = per_chunk * p->pt_chunks_max;
    std::vector<float> b(per_bin * F, 0.f);
    for (int f = 0; f < F; ++f)
        for (size_t j = 0; j < steps.size(); ++j)
            for (int part = 0; part < P; ++part)
                for (int e = 0; e < 2; ++e) {
                    const size_t hi_ = 2 * ((((size_t)part * K + steps[j].k) * 2 + e) * F + f);
                    const float hr = steps[j].sg * hnm[hi_], him = steps[j].sg * hnm[hi_ + 1];
                    for (int c = 0; c < 2; ++c)        // k = 2j + c
                        for (int ri = 0; ri < 2; ++ri) {  // n = ((part*2 + e)*2 + ri)
                            const float v = c == 0 ? (ri == 0 ? hr : him) : (ri == 0 ? -him : hr);
                            const int k = 2 * (int)j + c, n = (part * 2 + e) * 2 + ri;
                            const int h = k / KC, kl = k % KC;
                            const size_t off = ((size_t)(kl / 4) * (Npad / 8) + n / 8) * 32 + (n % 8) * 4 + kl % 4;
                            uint32_t u;
                            memcpy(&u, &v, 4);
                            u = (u + 0x1000u) & 0xffffe000u;  // round to nearest TF32
                            float hi;
                            memcpy(&hi, &u, 4);
                            float *blk = b.data() + f * per_bin + h * per_chunk;
                            blk[off] = hi;
                            blk[(size_t)KC * Npad + off] = v - hi;
                        }
                }
    return b;
}

void free_tables(rtb_sh_plan *p) {
    cudaFree(p->d_rmT); p->d_rmT = nullptr;
    cudaFree(p->d_zring); p->d_zring = nullptr;
    cudaFree(p->d_htab); p->d_htab = nullptr;
    cudaFree(p->d_signals); p->d_signals = nullptr;
    cudaFree(p->d_mg_htab); p->d_mg_htab = nullptr;
    cudaFree(p->d_mg_signals); p->d_mg_signals = nullptr;
    p->mg_order = -1;
    cudaFree(p->d_steps); p->d_steps = nullptr;
    cudaFree(p->d_tc_b); p->d_tc_b = nullptr;
    cudaFree(p->d_mma_r); p->d_mma_r = nullptr;
    p->mma_enabled = false;
    cudaFree(p->d_zspec); p->d_zspec = nullptr;
    cudaFree(p->d_qcur); p->d_qcur = nullptr;
    cudaFree(p->d_qlast); p->d_qlast = nullptr;
    cudaFree(p->d_pt_b); p->d_pt_b = nullptr;
    p->tables_set = false;
}

}  // namespace

extern "C" {

int rtb_sh_plan_create(rtb_sh_plan **out, int device, int n_streams, int block_length,
                       int n_channels, int sh_max_order, int n_ears, int n_partitions) {
    if (!out) return fail(RTB_ERR_INVALID, "plan output pointer is NULL");
    *out = nullptr;
    if (n_streams < 1) return fail(RTB_ERR_INVALID, "n_streams must be >= 1 (got %d)", n_streams);
    if (n_channels < 1) return fail(RTB_ERR_INVALID, "n_channels must be >= 1");
    if (sh_max_order < 0 || sh_max_order > RTB_MAX_ORDER)
        return fail(RTB_ERR_INVALID, "Invalid value \"%d\" for spherical harmonics order.", sh_max_order);
    if (n_ears != 2)
        return fail(RTB_ERR_UNSUPPORTED, "the SH renderer produces exactly 2 ear signals (got %d)", n_ears);
    if (!block_length_supported(block_length))
        return fail(RTB_ERR_UNSUPPORTED, "block length %d not supported by the SH renderer "
                    "(supported: " RTB_BLOCK_LENGTHS ")", block_length);
    if (n_partitions > 1 && block_length > 256)
        return fail(RTB_ERR_UNSUPPORTED, "the partitioned SH-domain HRIR extension supports block lengths up to 256 "
                    "(got %d with %d partitions)", block_length, n_partitions);
    if (n_partitions < 1) return fail(RTB_ERR_INVALID, "n_partitions must be >= 1");
    int ndev = 0;
    RTB_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(RTB_ERR_INVALID, "no CUDA device %d", device);

    rtb_sh_plan *p = new (std::nothrow) rtb_sh_plan();
    if (!p) return fail(RTB_ERR_NOMEM, "out of host memory");
    p->device = device;
    p->S = n_streams; p->B = block_length; p->C = n_channels; p->order = sh_max_order;
    p->K = (sh_max_order + 1) * (sh_max_order + 1);
    p->E = n_ears; p->P = n_partitions;
    p->F = block_length + 1; p->N2 = 2 * block_length;
    p->Fp = (p->F + 3) / 4 * 4;
    p->cur_order = p->last_order = sh_max_order;

    DeviceGuard guard(device);
    cudaError_t err = cudaStreamCreateWithFlags(&p->own_stream, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaMalloc(&p->d_tw, sizeof(float2) * p->N2);
    if (err == cudaSuccess) err = cudaMalloc(&p->d_win_in, sizeof(float) * p->B);
    if (err == cudaSuccess) err = cudaMalloc(&p->d_win_out, sizeof(float) * p->B);
    if (err == cudaSuccess) err = cudaMalloc(&p->d_radlast, sizeof(float) * 2 * p->S);
    if (err == cudaSuccess) err = cudaMemset(p->d_radlast, 0, sizeof(float) * 2 * p->S);
    if (err != cudaSuccess) {
        rtb_sh_plan_destroy(p);
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
    return RTB_OK;
}

int rtb_sh_plan_set_tables(rtb_sh_plan *p, const float *yw, const float *hnm, const int16_t *sh_m,
                           const int32_t *rev, const float *win_in, const float *win_out) {
    if (p) drop_graphs(p);
    if (!p || !yw || !hnm || !sh_m || !rev) return fail(RTB_ERR_INVALID, "NULL argument");
    DeviceGuard guard(p->device);
    const int K = p->K, C = p->C, F = p->F, B = p->B;
    auto YW = [&](int k, int c) { return make_float2(yw[2 * ((size_t)k * C + c)], yw[2 * ((size_t)k * C + c) + 1]); };

    // rev must be a permutation of 0..K-1
    std::vector<int> inv_rev(K, -1);
    for (int k = 0; k < K; ++k) {
        if (rev[k] < 0 || rev[k] >= K || inv_rev[rev[k]] >= 0)
            return fail(RTB_ERR_INVALID, "reverse index list is not a permutation");
        inv_rev[rev[k]] = k;
    }
    for (int k = 0; k < K; ++k)
        if (std::abs((int)sh_m[k]) > p->order)
            return fail(RTB_ERR_INVALID, "|sh_m[%d]| = %d exceeds the SH order %d", k, sh_m[k], p->order);

    // Does Y_w have the structure of complex spherical harmonics, rows ordered (n, m = -n..n)
    // with Y_w[(n,-m)] = (-1)^m conj(Y_w[(n,m)])?  True for both constructions of
    // FilterSetMiro.get_sh_configuration (_filter_set.py:1271-1278).
    bool standard = true;
    for (int n = 0, k = 0; n <= p->order && standard; ++n)
        for (int m = -n; m <= n; ++m, ++k)
            if (sh_m[k] != m || rev[k] != n * n + n - m) { standard = false; break; }
    bool symmetric = standard;
    if (standard) {
        double vmax = 0.0, dmax = 0.0;
        for (int k = 0; k < K; ++k) {
            const double sgn = (sh_m[k] & 1) ? -1.0 : 1.0;
            for (int c = 0; c < C; ++c) {
                const float2 a = YW(k, c), b = YW(rev[k], c);
                vmax = std::max(vmax, (double)std::max(std::fabs(a.x), std::fabs(a.y)));
                dmax = std::max(dmax, std::fabs(b.x - sgn * a.x));
                dmax = std::max(dmax, std::fabs(b.y + sgn * a.y));
            }
        }
        symmetric = dmax <= 1e-6 * vmax;
    }

    // complex FFT signals and the real analysis rows that feed them
    std::vector<ShSignal> sig;
    std::vector<int> sig_k1, sig_k2;   // H_nm row used by term 1 / term 2 (-1 = none)
    std::vector<float> sig_sign2;
    std::vector<std::vector<float>> rows;  // analysis rows [J][C]
    auto add_row = [&](int k, bool imag) {
        std::vector<float> r(C);
        for (int c = 0; c < C; ++c) r[c] = imag ? YW(k, c).y : YW(k, c).x;
        rows.push_back(r);
        return (int)rows.size() - 1;
    };
    if (symmetric) {
        for (int n = 0; n <= p->order; ++n)
            for (int m = 0; m <= n; ++m) {
                const int kpos = n * n + n + m, kneg = n * n + n - m;
                ShSignal s;
                s.row_re = add_row(kpos, false);
                s.row_im = m > 0 ? add_row(kpos, true) : -1;
                s.m1 = -m;                       // consumer k = (n,-m): S[rev k] = bin f
                s.m2 = m > 0 ? m : RTB_NO_TERM;  // consumer k = (n,+m): S[rev k] = (-1)^m conj(bin N2-f)
                sig.push_back(s);
                sig_k1.push_back(kneg);
                sig_k2.push_back(m > 0 ? kpos : -1);
                sig_sign2.push_back((m & 1) ? -1.f : 1.f);
            }
    } else {
        for (int k = 0; k < K; ++k) {  // signal k transforms S[rev[k]], consumed by H_nm[k]
            ShSignal s;
            s.row_re = add_row(rev[k], false);
            s.row_im = add_row(rev[k], true);
            s.m1 = sh_m[k];
            s.m2 = RTB_NO_TERM;
            sig.push_back(s);
            sig_k1.push_back(k);
            sig_k2.push_back(-1);
            sig_sign2.push_back(1.f);
        }
    }
    const int J = (int)rows.size(), NP = (int)sig.size();

    // K-A tiling: a warp tile has 4 row groups of JTH rows; choose JTH minimising padded work
    // plus a per-chunk overhead, then pack the analysis matrix as [chunk][C][4 groups x 8]
    int bestJT = 4, bestCost = 1 << 30;
    for (int jth = 4; jth <= 8; ++jth) {
        const int chunks = (J + 4 * jth - 1) / (4 * jth), cost = chunks * (4 * jth + 4);
        if (cost <= bestCost) { bestCost = cost; bestJT = jth; }
    }
    const int n_jchunks = (J + 4 * bestJT - 1) / (4 * bestJT), Jpad = n_jchunks * 4 * bestJT;
    std::vector<float> rmT((size_t)n_jchunks * C * 32, 0.f);
    for (int j = 0; j < J; ++j) {
        const int jc = j / (4 * bestJT), jr = (j % (4 * bestJT)) / bestJT, jj = j % bestJT;
        for (int c = 0; c < C; ++c) rmT[((size_t)jc * C + c) * 32 + jr * 8 + jj] = rows[j][c];
    }

    p->sig_k1 = sig_k1; p->sig_k2 = sig_k2; p->sig_sign2 = sig_sign2;
    std::vector<float4> htab = build_htab(p, hnm);

    // cross-fade windows (_convolver.py:741-747), pre-scaled by the inverse-FFT normalisation
    std::vector<float> wi(B), wo(B);
    for (int n = 0; n < B; ++n) {
        float o, i;
        if (win_in && win_out) { i = win_in[n]; o = win_out[n]; }
        else {
            const double c0 = cos(n * M_PI / (2.0 * (B - 1))), c1 = cos((B - 1 - n) * M_PI / (2.0 * (B - 1)));
            o = (float)(c0 * c0); i = (float)(c1 * c1);
        }
        wi[n] = i / (float)p->N2; wo[n] = o / (float)p->N2;  // N2 is a power of two: exact
    }

    free_tables(p);
    const size_t zbytes = sizeof(float) * 2 * (size_t)p->S * J * B;
    cudaError_t err = cudaMalloc(&p->d_rmT, sizeof(float) * rmT.size());
    if (err == cudaSuccess) err = cudaMalloc(&p->d_zring, zbytes);
    if (err == cudaSuccess) err = cudaMalloc(&p->d_htab, sizeof(float4) * htab.size());
    if (err == cudaSuccess) err = cudaMalloc(&p->d_signals, sizeof(ShSignal) * NP);
    if (p->P > 1) {
        // tensor-core partition MAC: default from 64 streams up (RTB_PART_TC=0 / 1 forces); needs
        // N = 4P <= 128 accumulator columns (two yaws, double buffered in 512 TMEM columns) and the
        // two phase tables next to four operand stages in shared memory
        p->pt_enabled = false;
        const int Npad = (4 * p->P + 15) / 16 * 16;
        if (const char *e2 = getenv("RTB_PT_OVERWRITE")) p->pt_overwrite = e2[0] != '0';
        if (const char *e3 = getenv("RTB_SPEC_SPLIT")) p->spec_split_forced = atoi(e3);
        const char *env = getenv("RTB_PART_TC");
        const bool want = env ? env[0] != '0' : p->S >= 64;
        const size_t stage_bytes = sizeof(float) * ((size_t)2 * 2 * RTB_PT_KC * 128 + (size_t)2 * RTB_PT_KC * Npad);
        const size_t pt_smem = RTB_PT_LG * stage_bytes + sizeof(float2) * 2 * 2 * (size_t)(p->order + 1) * 128;
        if (want && Npad <= 128 && pt_smem <= 220 * 1024) {
            p->pt_enabled = true;
            p->pt_Npad = Npad;
            p->pt_acc_cols = Npad <= 32 ? 32 : (Npad <= 64 ? 64 : 128);
            p->pt_Spad = ((long long)p->S + 127) / 128 * 128;
            p->pt_smem = (int)pt_smem;
            int n_steps = 0;
            for (int q = 0; q < NP; ++q) n_steps += sig[q].m2 != RTB_NO_TERM ? 2 : 1;
            p->pt_chunks_max = (2 * n_steps + RTB_PT_KC - 1) / RTB_PT_KC;
            int n_sm = 148;
            cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, p->device);
            p->pt_grid = n_sm;
        }
        const size_t zspec_elems = p->pt_enabled ? (size_t)p->N2 * NP * (size_t)p->pt_Spad : (size_t)p->S * NP * p->N2;
        if (err == cudaSuccess) err = cudaMalloc(&p->d_zspec, sizeof(float2) * zspec_elems);
        if (err == cudaSuccess) err = cudaMemset(p->d_zspec, 0, sizeof(float2) * zspec_elems);  // padding rows stay finite
        if (err == cudaSuccess) err = cudaMalloc(&p->d_qcur, sizeof(float2) * queue_elems(p));
        if (err == cudaSuccess) err = cudaMalloc(&p->d_qlast, sizeof(float2) * queue_elems(p));
        if (err == cudaSuccess) err = cudaMemset(p->d_qcur, 0, sizeof(float2) * queue_elems(p));
        if (err == cudaSuccess) err = cudaMemset(p->d_qlast, 0, sizeof(float2) * queue_elems(p));
    }
    if (err != cudaSuccess) {
        free_tables(p);
        return fail(err == cudaErrorMemoryAllocation ? RTB_ERR_NOMEM : RTB_ERR_CUDA,
                    "table allocation failed: %s", cudaGetErrorString(err));
    }
    RTB_CUDA(cudaMemcpy(p->d_rmT, rmT.data(), sizeof(float) * rmT.size(), cudaMemcpyHostToDevice));
    RTB_CUDA(cudaMemcpy(p->d_htab, htab.data(), sizeof(float4) * htab.size(), cudaMemcpyHostToDevice));
    RTB_CUDA(cudaMemcpy(p->d_signals, sig.data(), sizeof(ShSignal) * NP, cudaMemcpyHostToDevice));
    if (p->pt_enabled) {
        const std::vector<float> b = build_pt_b(p, hnm);
        cudaError_t e2 = cudaMalloc(&p->d_pt_b, sizeof(float) * b.size());
        if (e2 == cudaSuccess) e2 = cudaMemcpy(p->d_pt_b, b.data(), sizeof(float) * b.size(), cudaMemcpyHostToDevice);
        if (e2 != cudaSuccess) {
            free_tables(p);
            return fail(RTB_ERR_CUDA, "partition operand upload failed: %s", cudaGetErrorString(e2));
        }
    }
    if (p->P > 1) {
        int n_steps = 0;
        for (int q = 0; q < NP; ++q) n_steps += sig[q].m2 != RTB_NO_TERM ? 2 : 1;
        if (n_steps > RTB_MAC_MAX_STEPS) {
            free_tables(p);
            return fail(RTB_ERR_UNSUPPORTED, "the partitioned SH-domain HRIR extension supports up to %d "
                        "(signal, term) steps (SH order 12); this configuration has %d", RTB_MAC_MAX_STEPS, n_steps);
        }
    }
    {
        std::vector<ShStep> steps;
        p->steps_prefix.assign(NP + 1, 0);
        for (int q = 0; q < NP; ++q) {
            steps.push_back(ShStep{q, 0, sig[q].m1, 0});
            if (sig[q].m2 != RTB_NO_TERM) steps.push_back(ShStep{q, 1, sig[q].m2, 0});
            p->steps_prefix[q + 1] = (int)steps.size();
        }
        RTB_CUDA(cudaMalloc(&p->d_steps, sizeof(ShStep) * steps.size()));
        RTB_CUDA(cudaMemcpy(p->d_steps, steps.data(), sizeof(ShStep) * steps.size(), cudaMemcpyHostToDevice));
    }
    RTB_CUDA(cudaMemcpy(p->d_win_in, wi.data(), sizeof(float) * B, cudaMemcpyHostToDevice));
    RTB_CUDA(cudaMemcpy(p->d_win_out, wo.data(), sizeof(float) * B, cudaMemcpyHostToDevice));
    RTB_CUDA(cudaMemset(p->d_zring, 0, zbytes));
    RTB_CUDA(cudaMemset(p->d_radlast, 0, sizeof(float) * 2 * p->S));

    p->symmetric = symmetric;
    p->h_rows.assign((size_t)J * C, 0.f);
    for (int j = 0; j < J; ++j) std::copy(rows[j].begin(), rows[j].end(), p->h_rows.begin() + (size_t)j * C);
    p->h_sig = sig;
    p->h_shm.assign(sh_m, sh_m + K);
    p->mg_enabled = p->P == 1 && p->N2 <= 512;  // every warp-level transform size
    cudaDeviceGetAttribute(&p->n_sm, cudaDevAttrMultiProcessorCount, p->device);
    if (const char *env = getenv("RTB_RENDER_MG")) {
        p->mg_enabled = p->mg_enabled && env[0] != '0';
        p->mg_forced = env[0] == '1';
    }
    if (const char *env = getenv("RTB_PDL")) p->pdl = env[0] != '0';
    if (const char *env = getenv("RTB_GRAPH")) p->graphs_enabled = env[0] != '0';
    if (p->mg_enabled) {
        p->h_hnm.assign(hnm, hnm + 2 * (size_t)K * 2 * F);
        const int rc = build_mg_tables(p, p->order);
        if (rc != RTB_OK) return rc;
    }
    p->J = J; p->Jpad = Jpad; p->JT = bestJT; p->n_jchunks = n_jchunks; p->NP = NP;
    {   // bulk-copy pipeline of the analysis kernel: widest tile that leaves room for >= 2 stages
        const size_t rbytes = (size_t)n_jchunks * C * 32 * sizeof(float);
        p->an_rsmem = rbytes <= 48 * 1024;
        const size_t budget = 200 * 1024 - (p->an_rsmem ? rbytes : 0);
        p->an_TW = 0;
        for (int nrun = (B >= 64 ? 2 : 1); nrun >= 1 && !p->an_TW; --nrun) {
            p->an_nrun = nrun;
            for (int tw = B; tw >= 32 * nrun; tw /= 2)
                if (2 * (size_t)C * tw * sizeof(float) <= budget) { p->an_TW = tw; break; }
        }
        if (!p->an_TW) p->an_TW = 32;
        const size_t tile = (size_t)C * p->an_TW * sizeof(float);
        if (2 * tile > budget)
            return fail(RTB_ERR_UNSUPPORTED, "%d array channels do not fit the analysis kernel's tiles", C);
        size_t stage_budget = 68 * 1024;  // three CTAs per SM when the tiles are small
        if (const char *env = getenv("RTB_AN_STAGE_KB")) stage_budget = (size_t)atoi(env) * 1024;
        p->an_stages = (int)std::min<size_t>(RTB_AN_MAX_STAGES, std::max<size_t>(2, stage_budget / tile));
        p->an_smem = (int)(p->an_stages * tile + (p->an_rsmem ? rbytes : 0));
        int n_sm = 148;
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, p->device);
        const int ctas_per_sm = std::max(1, std::min(4, (int)((224 * 1024) / (p->an_smem + 2048))));
        p->an_grid = n_sm * ctas_per_sm;  // persistent CTAs; capped by the tile count at launch
        // tensor-core variant: wide arrays (FP32-bound on SIMT), whole analysis matrix resident in smem
        p->tc_enabled = false;
        const char *env = getenv("RTB_ANALYSIS_TC");
        const bool want = env ? (env[0] != '0') : (C >= 32);
        // K is cut into pipeline chunks of RTB_TC_KC = 32 channels (one channel quad per loader warp;
        // the last chunk issues only the k-steps that hold channels).  If the split analysis matrix
        // fits next to two operand stages it stays resident, otherwise its K-chunks are streamed
        // through the stages.  The operand ring takes the stages that fit (2 - 4, RTB_TC_NST caps it).
        const int Npad = (J + 15) / 16 * 16;
        const int KC = RTB_TC_KC, Kpad = (C + KC - 1) / KC * KC, Kres = (C + 7) / 8 * 8;
        int nst = 0;
        size_t tc_smem = 0;
        bool stream_b = false;
        const size_t tc_budget = 216 * 1024;
        int max_nst = RTB_TC_MAX_ST;
        if (const char *e5 = getenv("RTB_TC_NST")) max_nst = std::max(2, std::min(RTB_TC_MAX_ST, atoi(e5)));
        {
            const size_t a_stage = sizeof(float) * 2 * KC * RTB_TC_M;       // hi + lo
            const size_t b_res = sizeof(float) * 2 * (size_t)Kres * Npad;   // hi + lo
            const size_t b_stage = sizeof(float) * 2 * (size_t)KC * Npad;
            if (b_res + 2 * a_stage <= tc_budget) {
                nst = (int)std::min<size_t>(max_nst, (tc_budget - b_res) / a_stage) >= 4 ? 4 : 2;  // one loader group per stage
                tc_smem = b_res + nst * a_stage;
            } else if (2 * (a_stage + b_stage) <= tc_budget) {
                stream_b = true;
                nst = (int)std::min<size_t>(max_nst, tc_budget / (a_stage + b_stage)) >= 4 ? 4 : 2;
                tc_smem = nst * (a_stage + b_stage);
            }
        }
        if (want && nst >= 2 && Npad <= 256 && B >= 32 && (B & (B - 1)) == 0) {
            std::vector<float> bsrc(2 * (size_t)Kpad * Npad, 0.f);
            for (int j = 0; j < J; ++j)
                for (int c = 0; c < C; ++c) {
                    const float v = rows[j][c];
                    uint32_t u;
                    memcpy(&u, &v, 4);
                    u = (u + 0x1000u) & 0xffffe000u;  // round to nearest TF32
                    float hi;
                    memcpy(&hi, &u, 4);
                    const size_t off = (((size_t)(c / 4) * (Npad / 8) + j / 8) * 8 + j % 8) * 4 + c % 4;
                    bsrc[off] = hi;
                    bsrc[(size_t)Kpad * Npad + off] = v - hi;
                }
            cudaError_t e2 = cudaMalloc(&p->d_tc_b, sizeof(float) * bsrc.size());
            if (e2 == cudaSuccess)
                e2 = cudaMemcpy(p->d_tc_b, bsrc.data(), sizeof(float) * bsrc.size(), cudaMemcpyHostToDevice);
            if (e2 != cudaSuccess) return fail(RTB_ERR_CUDA, "tensor-core table upload failed: %s", cudaGetErrorString(e2));
            p->tc_enabled = true;
            p->tc_stream_b = stream_b;
            p->tc_nst = nst;
            p->tc_Kpad = Kpad; p->tc_Npad = Npad; p->tc_KC = KC; p->tc_smem = (int)tc_smem;
            p->tc_cols = Npad <= 32 ? 32 : (Npad <= 64 ? 64 : (Npad <= 128 ? 128 : 256));  // per accumulator
            p->tc_grid = n_sm;
        }
    }
    {   // warp-level tensor-core variant for narrow arrays (RTB_ANALYSIS_MMA=0 disables)
        const char *env = getenv("RTB_ANALYSIS_MMA");
        const bool want = env ? env[0] != '0' : true;
        if (want && !p->tc_enabled && C <= RTB_MMA_CP && B >= 64) {
            const int njc = (J + RTB_MMA_JP - 1) / RTB_MMA_JP;
            std::vector<float> r((size_t)njc * 2 * RTB_MMA_JP * RTB_MMA_CP, 0.f);
            for (int j = 0; j < J; ++j)
                for (int c = 0; c < C; ++c) {
                    const float v = rows[j][c];
                    uint32_t u;
                    memcpy(&u, &v, 4);
                    u = (u + 0x1000u) & 0xffffe000u;  // round to nearest TF32
                    float hi;
                    memcpy(&hi, &u, 4);
                    const size_t o = ((size_t)(j / RTB_MMA_JP) * 2 * RTB_MMA_JP + j % RTB_MMA_JP) * RTB_MMA_CP + c;
                    r[o] = hi;
                    r[o + (size_t)RTB_MMA_JP * RTB_MMA_CP] = v - hi;
                }
            cudaError_t e2 = cudaMalloc(&p->d_mma_r, sizeof(float) * r.size());
            if (e2 == cudaSuccess) e2 = cudaMemcpy(p->d_mma_r, r.data(), sizeof(float) * r.size(), cudaMemcpyHostToDevice);
            if (e2 != cudaSuccess) return fail(RTB_ERR_CUDA, "analysis table upload failed: %s", cudaGetErrorString(e2));
            p->mma_enabled = true;
            p->mma_jchunks = njc;
            p->mma_TW = std::min(B, RTB_MMA_TW);
            if (const char *e4 = getenv("RTB_MMA_TW")) p->mma_TW = std::max(64, std::min(p->mma_TW, atoi(e4)));
            p->mma_stages = 2;
            if (const char *e3 = getenv("RTB_MMA_STAGES")) p->mma_stages = std::max(2, std::min(RTB_AN_MAX_STAGES, atoi(e3)));
            p->mma_smem = p->mma_stages * RTB_MMA_CP * (p->mma_TW + RTB_MMA_PAD) * (int)sizeof(float);
            int n_sm = 148;
            cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, p->device);
            p->mma_grid = n_sm * std::max(1, std::min(3, (224 * 1024) / (p->mma_smem + 2048)));
        }
    }
    p->slot = 0; p->rad_idx = 0;
    p->head_cur = p->head_last = 0;
    p->pt_stale_cur = p->pt_stale_last = -1;
    p->cur_order = p->last_order = p->order;
    p->last_valid = false;
    p->tables_set = true;
    RTB_CUDA(cudaDeviceSynchronize());  // table uploads / state memsets done before a non-blocking stream runs
    return RTB_OK;
}

int rtb_sh_plan_update_filter(rtb_sh_plan *p, const float *hnm) {
    if (p) drop_graphs(p);
    if (!p || !hnm) return fail(RTB_ERR_INVALID, "NULL argument");
    if (!p->tables_set)
        return fail(RTB_ERR_STATE, "blocks in spherical harmonics domain have not been calculated yet.");
    DeviceGuard guard(p->device);
    std::vector<float4> htab = build_htab(p, hnm);
    // the renderer may be running on another stream: plain (synchronising) copy, like the
    // reference's in-place `_irs_blocks_nm *= comp_nm` under a cleared IS_RUNNING
    RTB_CUDA(cudaDeviceSynchronize());
    RTB_CUDA(cudaMemcpy(p->d_htab, htab.data(), sizeof(float4) * htab.size(), cudaMemcpyHostToDevice));
    if (p->pt_enabled) {
        const std::vector<float> b = build_pt_b(p, hnm);
        RTB_CUDA(cudaMemcpy(p->d_pt_b, b.data(), sizeof(float) * b.size(), cudaMemcpyHostToDevice));
    }
    if (p->mg_enabled) {
        p->h_hnm.assign(hnm, hnm + 2 * (size_t)p->K * 2 * p->F);
        return build_mg_tables(p, p->mg_order >= 0 ? p->mg_order : p->cur_order);
    }
    return RTB_OK;
}

int rtb_sh_plan_set_source(rtb_sh_plan *p, float source_azim_deg, int tracker_dir) {
    if (p) drop_graphs(p);
    if (!p) return fail(RTB_ERR_INVALID, "NULL plan");
    if (tracker_dir != 1 && tracker_dir != -1) return fail(RTB_ERR_INVALID, "tracker_dir must be +1 or -1");
    p->src_azim16 = round_to_half(source_azim_deg);
    p->tracker_dir = (float)tracker_dir;
    return RTB_OK;
}

int rtb_sh_plan_set_order(rtb_sh_plan *p, int order) {
    if (!p) return fail(RTB_ERR_INVALID, "NULL plan");
    if (order < 0 || order > p->order)
        return fail(RTB_ERR_INVALID, "value \"%d\" out of bounds, \"set SH processing order\" ignored.", order);
    p->cur_order = order;
    return RTB_OK;
}

int rtb_sh_plan_set_passthrough(rtb_sh_plan *p, int on) {
    if (!p) return fail(RTB_ERR_INVALID, "NULL plan");
    p->passthrough = on != 0;
    return RTB_OK;
}

int rtb_sh_plan_set_crossfade(rtb_sh_plan *p, int on) {
    if (!p) return fail(RTB_ERR_INVALID, "NULL plan");
    p->crossfade = on != 0;
    if (!p->crossfade) {
        p->last_valid = false;  // _last_sh_azim_nm.fill(0), _convolver.py:1338-1339
        if (p->d_qlast) {       // _last_blocks_fd.fill(0), _convolver.py:887-889
            DeviceGuard guard(p->device);
            RTB_CUDA(cudaDeviceSynchronize());
            RTB_CUDA(cudaMemset(p->d_qlast, 0, sizeof(float2) * queue_elems(p)));
            RTB_CUDA(cudaDeviceSynchronize());
            p->pt_stale_last = -1;
        }
    }
    return RTB_OK;
}

int rtb_sh_plan_set_profiling(rtb_sh_plan *p, int on) {
    if (!p) return fail(RTB_ERR_INVALID, "NULL plan");
    DeviceGuard guard(p->device);
    if (on && !p->ev[0])
        for (auto &e : p->ev) RTB_CUDA(cudaEventCreate(&e));
    p->profiling = on != 0;
    p->ev_valid = false;
    return RTB_OK;
}

int rtb_sh_plan_reset(rtb_sh_plan *p) {
    if (p) drop_graphs(p);
    if (!p) return fail(RTB_ERR_INVALID, "NULL plan");
    if (!p->tables_set) return RTB_OK;
    DeviceGuard guard(p->device);
    // _input_block_td.fill(0); the P == 1 spectra buffers are rewritten every block anyway.  Legacy-stream
    // memsets do not order against non-blocking streams: drain the blocks in flight, then the memsets.
    RTB_CUDA(cudaDeviceSynchronize());
    RTB_CUDA(cudaMemset(p->d_zring, 0, sizeof(float) * 2 * (size_t)p->S * p->J * p->B));
    if (p->P > 1) {  // _blocks_fd.fill(0); _last_blocks_fd.fill(0)
        RTB_CUDA(cudaMemset(p->d_qcur, 0, sizeof(float2) * queue_elems(p)));
        RTB_CUDA(cudaMemset(p->d_qlast, 0, sizeof(float2) * queue_elems(p)));
        p->pt_stale_cur = p->pt_stale_last = -1;
    }
    if (p->hp_P > 0) {  // the HPCF client's own _clear_buffers (_convolver.py:500-505)
        RTB_CUDA(cudaMemset(p->d_hp_x, 0, sizeof(float) * 2 * (size_t)p->S * 2 * p->B));
        RTB_CUDA(cudaMemset(p->d_hp_hist, 0, sizeof(float2) * (size_t)p->S * p->hp_P * 2 * p->Fp));
    }
    RTB_CUDA(cudaDeviceSynchronize());
    return RTB_OK;
}

// Attach (or, with n_partitions == 0, remove) a 2-channel headphone-compensation filter behind the
// renderer: hfd [P][2][F] complex = FilterSet.get_filter_blocks_fd() of the HPCF set
// (_filter_set.py:625-664).  Same result as the reference's separate HPCF client
// (OverlapSaveConvolver.filter_block, _convolver.py:524-571) fed with the renderer's output.
int rtb_sh_plan_set_hpcf(rtb_sh_plan *p, const float *hfd, int n_partitions) {
    if (p) drop_graphs(p);
    if (!p) return fail(RTB_ERR_INVALID, "NULL plan");
    if (n_partitions < 0 || (n_partitions > 0 && !hfd)) return fail(RTB_ERR_INVALID, "bad HPCF argument");
    DeviceGuard guard(p->device);
    RTB_CUDA(cudaDeviceSynchronize());
    cudaFree(p->d_hp_hfd); cudaFree(p->d_hp_hist); cudaFree(p->d_hp_x);
    p->d_hp_hfd = p->d_hp_hist = nullptr; p->d_hp_x = nullptr;
    p->hp_P = 0; p->hp_slot = 0; p->hp_xidx = 0;
    if (n_partitions == 0) return RTB_OK;
    const int F = p->F, Fp = p->Fp, P = n_partitions;
    std::vector<float2> h((size_t)P * 2 * Fp, make_float2(0.f, 0.f));
    for (int q = 0; q < P; ++q)
        for (int e = 0; e < 2; ++e)
            for (int f = 0; f < F; ++f) {
                const size_t i = 2 * (((size_t)q * 2 + e) * F + f);
                h[((size_t)q * 2 + e) * Fp + f] = make_float2(hfd[i], hfd[i + 1]);
            }
    const size_t xbytes = sizeof(float) * 2 * (size_t)p->S * 2 * p->B;
    const size_t hbytes = sizeof(float2) * (size_t)p->S * P * 2 * Fp;
    cudaError_t err = cudaMalloc(&p->d_hp_hfd, sizeof(float2) * h.size());
    if (err == cudaSuccess) err = cudaMalloc(&p->d_hp_hist, hbytes);
    if (err == cudaSuccess) err = cudaMalloc(&p->d_hp_x, xbytes);
    if (err == cudaSuccess) err = cudaMemset(p->d_hp_hist, 0, hbytes);
    if (err == cudaSuccess) err = cudaMemset(p->d_hp_x, 0, xbytes);
    if (err == cudaSuccess) err = cudaMemcpy(p->d_hp_hfd, h.data(), sizeof(float2) * h.size(), cudaMemcpyHostToDevice);
    if (err == cudaSuccess) err = cudaDeviceSynchronize();
    if (err != cudaSuccess) {
        cudaFree(p->d_hp_hfd); cudaFree(p->d_hp_hist); cudaFree(p->d_hp_x);
        p->d_hp_hfd = p->d_hp_hist = nullptr; p->d_hp_x = nullptr;
        return fail(err == cudaErrorMemoryAllocation ? RTB_ERR_NOMEM : RTB_ERR_CUDA,
                    "HPCF allocation failed: %s", cudaGetErrorString(err));
    }
    p->hp_P = P;
    return RTB_OK;
}

namespace {

// stream-minor queues [F][P][2][Spad]: clear the slots whose consumed heads were left for the next
// partition MAC to overwrite (a block that does not run that kernel follows)
void clear_pt_stale(rtb_sh_plan *p, cudaStream_t st) {
    const size_t width = sizeof(float2) * 2 * (size_t)p->pt_Spad, pitch = width * p->P;
    if (p->pt_stale_cur >= 0)
        cudaMemset2DAsync(p->d_qcur + (size_t)p->pt_stale_cur * 2 * p->pt_Spad, pitch, 0, width, p->F, st);
    if (p->pt_stale_last >= 0)
        cudaMemset2DAsync(p->d_qlast + (size_t)p->pt_stale_last * 2 * p->pt_Spad, pitch, 0, width, p->F, st);
    p->pt_stale_cur = p->pt_stale_last = -1;
}

// kernels for streams [s0, s0 + count) of the current block; in/yaw/out point at stream s0
void launch_range(rtb_sh_plan *p, int s0, int count, const float *d_in, const float *d_yaw_deg,
                  float *d_out, cudaStream_t st, bool rows_given = false) {
    const size_t zhalf = (size_t)p->S * p->J * p->B, zo = (size_t)s0 * p->J * p->B;
    float *zcur = p->d_zring + (size_t)p->slot * zhalf + zo;
    const float *zprev = p->d_zring + (size_t)(p->slot ^ 1) * zhalf + zo;
    if (!rows_given) {  // (pre-rendered mode: the rows of this block were written by the pre-renderer)
        launch_analysis(p, d_in, zcur, count, st);  // a1/a2: the overlap-save state advances in every mode
        p->launches++;
    }
    // with an HPCF attached the renderer's binaural block goes to the filter's overlap-save state and
    // the filter writes what the caller receives (fused into the render kernels, stand-alone otherwise)
    HpcfParams hp = {};
    float *const d_out_final = d_out;
    if (p->hp_P > 0) {
        const size_t xs = (size_t)p->S * 2 * p->B;
        hp.x_prev = p->d_hp_x + (size_t)(p->hp_xidx ^ 1) * xs + (size_t)s0 * 2 * p->B;
        hp.x_cur = p->d_hp_x + (size_t)p->hp_xidx * xs + (size_t)s0 * 2 * p->B;
        hp.hist = p->d_hp_hist + (size_t)s0 * p->hp_P * 2 * p->Fp;
        hp.hfd = p->d_hp_hfd;
        hp.out = d_out_final;
        hp.P = p->hp_P; hp.Fp = p->Fp; hp.slot = p->hp_slot;
        d_out = hp.x_cur;
    }
    bool hp_fused = false;
    if (p->profiling && s0 == 0) cudaEventRecord(p->ev[1], st);
    if (p->passthrough) {
        p->last_render_kernel = 4;
        const long long total = (long long)count * p->E * p->B;
        sh_passthrough_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_in, d_out, count, p->C, p->E, p->B);
        if (p->P > 1 && p->pt_enabled) {
            clear_pt_stale(p, st);
            const long long n = (long long)p->F * 2 * count;
            sh_queue_clear_head_soa_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
                p->d_qcur, count, p->F, p->P, p->pt_Spad, p->head_cur);
            p->launches++;
        } else if (p->P > 1) {  // queue slot 0 is overwritten and leaves the queue (_convolver.py:558-562, :653-659)
            const long long n = (long long)count * 2 * p->Fp;
            sh_queue_clear_head_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
                p->d_qcur + (size_t)s0 * p->P * 2 * p->Fp, count, p->P, p->Fp, p->head_cur);
            p->launches++;
        }
    } else if (p->P > 1) {
        p->last_render_kernel = 3;
        ShPartParams prm;
        prm.zprev = zprev; prm.zcur = zcur;
        prm.zspec = p->d_zspec + (size_t)s0 * p->NP * p->N2;
        prm.htab = p->d_htab; prm.signals = p->d_signals; prm.tw = p->d_tw;
        prm.win_in = p->d_win_in; prm.win_out = p->d_win_out;
        prm.yaw_deg = d_yaw_deg;
        prm.rad_last_in = p->d_radlast + (size_t)p->rad_idx * p->S + s0;
        prm.rad_last_out = p->d_radlast + (size_t)(p->rad_idx ^ 1) * p->S + s0;
        prm.qcur = p->d_qcur + (size_t)s0 * p->P * 2 * p->Fp;
        prm.qlast = p->d_qlast + (size_t)s0 * p->P * 2 * p->Fp;
        prm.out = d_out;
        prm.S = count; prm.J = p->J; prm.P = p->P; prm.Fp = p->Fp; prm.NPT = p->NP;
        prm.np_cur = signals_for_order(p, p->cur_order);
        prm.np_last = (p->crossfade && p->last_valid) ? signals_for_order(p, p->last_order) : 0;
        prm.order_max = p->order;
        prm.crossfade = p->crossfade ? 1 : 0;
        prm.head_cur = p->head_cur; prm.head_last = p->head_last;
        prm.steps = p->d_steps;
        prm.nsteps_cur = p->steps_prefix[prm.np_cur];
        prm.nsteps_last = p->steps_prefix[prm.np_last];
        prm.tracker_dir = p->tracker_dir; prm.src_azim16 = p->src_azim16;
        const unsigned gw = (unsigned)((count + 3) / 4);
        const dim3 gmac2((unsigned)((p->F + 31) / 32), (unsigned)((count + 2 * RTB_MAC_WARPS - 1) / (2 * RTB_MAC_WARPS)));
        if (p->pt_enabled) {  // tensor-core partition MAC on the stream-minor layout (whole plan: s0 == 0)
            p->last_render_kernel = 6;
            const unsigned gs = (unsigned)((count + RTB_SPEC_NW - 1) / RTB_SPEC_NW);
            // shares of the signal rounds per stream group (grid y): enough CTAs for >= 8 waves of the
            // three resident CTAs per SM, at least 6 rounds per CTA (RTB_SPEC_SPLIT overrides)
            const int np_spec = std::max(prm.np_cur, prm.np_last);
            int spec_split = (int)std::min<long long>(std::max(1, np_spec / 6),
                                                      (8LL * 3 * p->n_sm + gs - 1) / gs);
            if (p->spec_split_forced > 0) spec_split = p->spec_split_forced;
            spec_split = std::max(1, std::min(spec_split, std::max(1, np_spec)));
#define RTB_PT_LAUNCH(N)                                                                                     \
    {                                                                                                        \
        const size_t sm = sizeof(float2) * (size_t)(FftCfg<N>::GROUPS * N) * (RTB_SPEC_NW + 1);              \
        kernel_attrs_once(sh_spectra_soa_kernel<N>, 100 * 1024, true);                                       \
        sh_spectra_soa_kernel<N><<<dim3(gs, spec_split), 32 * RTB_SPEC_NW, sm, st>>>(prm, p->d_zspec, p->pt_Spad); \
    }
            switch (p->N2) {
                case 64: RTB_PT_LAUNCH(64) break;
                case 128: RTB_PT_LAUNCH(128) break;
                case 256: RTB_PT_LAUNCH(256) break;
                default: RTB_PT_LAUNCH(512) break;
            }
#undef RTB_PT_LAUNCH
            ShPartTcParams tp;
            tp.zspecT = p->d_zspec; tp.bsrc = p->d_pt_b; tp.steps = p->d_steps;
            tp.yaw_deg = d_yaw_deg; tp.rad_last_in = prm.rad_last_in;
            tp.qcurT = p->d_qcur; tp.qlastT = p->d_qlast;
            tp.Spad = p->pt_Spad; tp.S = count;
            tp.F = p->F; tp.N2 = p->N2; tp.NPT = p->NP; tp.P = p->P; tp.Npad = p->pt_Npad; tp.acc_cols = p->pt_acc_cols;
            tp.nsteps_cur = prm.nsteps_cur; tp.nsteps_last = prm.nsteps_last;
            const int nsteps = std::max(prm.nsteps_cur, prm.nsteps_last);
            tp.n_chunks = (2 * nsteps + RTB_PT_KC - 1) / RTB_PT_KC;
            tp.n_chunks_max = p->pt_chunks_max;
            tp.order_max = p->order; tp.do_last = p->crossfade ? 1 : 0;
            tp.head_cur = p->head_cur; tp.head_last = p->head_last;
            tp.overwrite_tail = (tp.n_chunks > 0 && p->pt_overwrite) ? 1 : 0;
            if (!tp.overwrite_tail) clear_pt_stale(p, st);
            tp.tracker_dir = p->tracker_dir; tp.src_azim16 = p->src_azim16;
            kernel_attrs_once(sh_partition_tc_kernel, 220 * 1024);  // + 3 KB static (barriers, step list)
            const int n_stiles_total = (count + 127) / 128;
            if (tp.n_chunks > 0)
                // a launch covers at most as many stream tiles as there are CTAs, so that the contiguous
                // run of tiles of a CTA (<= F tiles) meets at most two stream tiles
                for (int st0 = 0, per = std::min(RTB_PT_MAX_STILES_PER_LAUNCH, p->pt_grid); st0 < n_stiles_total; st0 += per) {
                    tp.s0 = st0 * 128;
                    tp.n_stiles = std::min(per, n_stiles_total - st0);
                    ShPartTcParams q = tp;
                    q.zspecT = tp.zspecT + tp.s0; q.qcurT = tp.qcurT + tp.s0; q.qlastT = tp.qlastT + tp.s0;
                    const long long n_tiles = (long long)q.n_stiles * p->F;
                    const int grid = (int)std::min<long long>(n_tiles, p->pt_grid);
                    sh_partition_tc_kernel<<<grid, RTB_PT_THREADS, p->pt_smem, st>>>(q);
                    p->launches++;
                }
#define RTB_PT_TAIL(N) sh_partition_tail_soa_kernel<N><<<gw, 128, 0, st>>>(prm, p->d_qcur, p->d_qlast, p->pt_Spad, tp.overwrite_tail ? 0 : 1);
            switch (p->N2) {
                case 64: RTB_PT_TAIL(64) break;
                case 128: RTB_PT_TAIL(128) break;
                case 256: RTB_PT_TAIL(256) break;
                default: RTB_PT_TAIL(512) break;
            }
#undef RTB_PT_TAIL
            if (tp.overwrite_tail) {  // the consumed heads stay; the next block's MAC overwrites them
                p->pt_stale_cur = p->head_cur;
                if (p->crossfade) p->pt_stale_last = p->head_last;
            }
            p->launches++;
        } else {
#define RTB_PART_LAUNCH(N)                                                                  \
    case N:                                                                                 \
        sh_spectra_kernel<N><<<gw, 128, 0, st>>>(prm);                                      \
        sh_partition_mac2_kernel<N><<<gmac2, 32 * RTB_MAC_WARPS, 0, st>>>(prm);             \
        sh_partition_tail_kernel<N><<<gw, 128, 0, st>>>(prm);                               \
        break;
        switch (p->N2) {
            RTB_PART_LAUNCH(64) RTB_PART_LAUNCH(128) RTB_PART_LAUNCH(256)
            default: RTB_PART_LAUNCH(512)
        }
#undef RTB_PART_LAUNCH
        p->launches += 2;
        }
    } else {
        const int np_cur = signals_for_order(p, p->cur_order);
        const int np_last = (p->crossfade && p->last_valid) ? signals_for_order(p, p->last_order) : 0;
        // steady state (both yaws at the same order): the m-grouped kernel; its tables follow
        // the current order (rebuilt on the first block after the order settled)
        // several streams share a warp below 128-sample blocks: 8 / 16 streams per CTA under-fill the GPU
        // for small batches, where the term-wise kernel (one stream per warp) is faster (measured, c4:
        // 1 024 streams 93 vs 86 us, 4 096 streams 0.747 vs 0.804 ms for the chain)
        // (RTB_RENDER_MG=1 at plan creation forces the m-grouped kernel)
        bool use_mg = p->mg_enabled && (np_last == 0 || np_last == np_cur) &&
                      (p->N2 >= 256 || count >= 2048 || p->mg_forced);
        if (use_mg && p->mg_order != p->cur_order) {
            cudaStreamSynchronize(st);
            use_mg = build_mg_tables(p, p->cur_order) == RTB_OK && p->mg_enabled;
        }
        p->last_render_kernel = use_mg ? 1 : 0;
        if (use_mg) {
            ShRenderMgParams prm;
            prm.zprev = zprev; prm.zcur = zcur;
            prm.htab = p->d_mg_htab; prm.signals = p->d_mg_signals; prm.tw = p->d_tw;
            prm.win_in = p->d_win_in; prm.win_out = p->d_win_out;
            prm.yaw_deg = d_yaw_deg;
            prm.rad_last_in = p->d_radlast + (size_t)p->rad_idx * p->S + s0;
            prm.rad_last_out = p->d_radlast + (size_t)(p->rad_idx ^ 1) * p->S + s0;
            prm.out = d_out;
            prm.S = count; prm.J = p->J; prm.B = p->B;
            prm.np = p->mg_np;
            prm.last_on = np_last > 0 ? 1 : 0;
            prm.order_max = p->order;
            prm.crossfade = p->crossfade ? 1 : 0;
            prm.tracker_dir = p->tracker_dir; prm.src_azim16 = p->src_azim16;
            prm.hp = hp; hp_fused = true;
            launch_render_mg(p, prm, st);
        } else {
            ShRenderParams prm;
            prm.zprev = zprev; prm.zcur = zcur;
            prm.htab = p->d_htab; prm.signals = p->d_signals; prm.tw = p->d_tw;
            prm.win_in = p->d_win_in; prm.win_out = p->d_win_out;
            prm.yaw_deg = d_yaw_deg;
            prm.rad_last_in = p->d_radlast + (size_t)p->rad_idx * p->S + s0;
            prm.rad_last_out = p->d_radlast + (size_t)(p->rad_idx ^ 1) * p->S + s0;
            prm.out = d_out;
            prm.S = count; prm.J = p->J; prm.B = p->B;
            prm.np_cur = np_cur;
            prm.np_last = np_last;
            prm.order_max = p->order;
            prm.crossfade = p->crossfade ? 1 : 0;
            prm.tracker_dir = p->tracker_dir; prm.src_azim16 = p->src_azim16;
            prm.hp = hp; hp_fused = true;
            launch_render(p, prm, st);
        }
    }
    p->launches++;
    if (p->hp_P > 0 && !hp_fused) {  // passthrough / partitioned paths: the filter as its own launch
#define RTB_HP_LAUNCH(N) fft_kernel_launch<N>(hpcf_kernel<N>, hp, count, st, 0, 0, p->d_tw, count)
        RTB_DISPATCH_N2(p->N2, RTB_HP_LAUNCH)
#undef RTB_HP_LAUNCH
        p->launches++;
    }
}

// block boundary: flip the ping-pong state once all ranges of the block are enqueued
void advance_block(rtb_sh_plan *p) {
    if (!p->passthrough && p->crossfade) {  // _last_sh_azim_nm = sh_azim_nm (_convolver.py:1311)
        p->rad_idx ^= 1;
        p->last_valid = true;
        p->last_order = p->cur_order;
        if (p->P > 1) p->head_last = (p->head_last + 1) % p->P;  // the last-yaw queue rolled too
    }
    if (p->P > 1) p->head_cur = (p->head_cur + 1) % p->P;
    p->slot ^= 1;
    if (p->hp_P > 0) {
        p->hp_xidx ^= 1;
        p->hp_slot = (p->hp_slot + 1) % p->hp_P;
    }
}

}  // namespace

int rtb_sh_process_block(rtb_sh_plan *p, const float *d_in, const float *d_yaw_deg, float *d_out,
                         void *cuda_stream) {
    if (!p || !d_in || !d_out) return fail(RTB_ERR_INVALID, "NULL argument");
    if ((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_out)) & 15)
        return fail(RTB_ERR_INVALID, "device buffers must be 16-byte aligned (bulk copies, 16-byte loads and stores)");
    if (!p->tables_set)
        return fail(RTB_ERR_STATE, "blocks in spherical harmonics domain have not been calculated yet.");
    if (!d_yaw_deg && !p->passthrough) return fail(RTB_ERR_INVALID, "NULL yaw pointer");
    DeviceGuard guard(p->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (p->profiling) cudaEventRecord(p->ev[0], st);
    launch_range(p, 0, p->S, d_in, d_yaw_deg, d_out, st);
    if (p->profiling) { cudaEventRecord(p->ev[2], st); p->ev_valid = true; }
    advance_block(p);
    RTB_CUDA(cudaGetLastError());
    return RTB_OK;
}

// ---- pre-rendered mode -------------------------------------------------------------------------
// The array signals enter the chain only through the analysis rows z = R x (R = Re / Im of Y_w,
// rtb_sh_plan_get_analysis_rows).  When they are themselves the output of a linear time-invariant
// pre-renderer -- the reference's ARIR client, a mono source convolved with one room impulse
// response per array channel (_run.py:240-288, OverlapSaveConvolver with a mono input) -- R can be
// folded into its filters once at set-up time: z_j = (sum_c R[j][c] arir_c) * source.  The
// pre-renderer then produces the J analysis rows directly (J = 169 instead of 434 channels at order
// 12 on a 434-channel array), writes them into the plan's row buffer, and the renderer skips its
// analysis GEMM.  Same result up to fp32 reassociation.
int rtb_sh_plan_get_analysis_rows(rtb_sh_plan *p, float *rows) {
    if (!p || !rows) return fail(RTB_ERR_INVALID, "NULL argument");
    if (!p->tables_set)
        return fail(RTB_ERR_STATE, "blocks in spherical harmonics domain have not been calculated yet.");
    std::copy(p->h_rows.begin(), p->h_rows.end(), rows);
    return RTB_OK;
}

int rtb_sh_plan_rows_buffer(rtb_sh_plan *p, float **d_rows) {
    if (!p || !d_rows) return fail(RTB_ERR_INVALID, "NULL argument");
    if (!p->tables_set)
        return fail(RTB_ERR_STATE, "blocks in spherical harmonics domain have not been calculated yet.");
    *d_rows = p->d_zring + (size_t)p->slot * ((size_t)p->S * p->J * p->B);
    return RTB_OK;
}

int rtb_sh_process_block_rows(rtb_sh_plan *p, const float *d_yaw_deg, float *d_out, void *cuda_stream) {
    if (!p || !d_out || !d_yaw_deg) return fail(RTB_ERR_INVALID, "NULL argument");
    if (reinterpret_cast<uintptr_t>(d_out) & 15) return fail(RTB_ERR_INVALID, "device buffers must be 16-byte aligned");
    if (!p->tables_set)
        return fail(RTB_ERR_STATE, "blocks in spherical harmonics domain have not been calculated yet.");
    if (p->passthrough)
        return fail(RTB_ERR_UNSUPPORTED, "passthrough copies array channels to the ears; there are none in pre-rendered mode");
    DeviceGuard guard(p->device);
    launch_range(p, 0, p->S, nullptr, d_yaw_deg, d_out, (cudaStream_t)cuda_stream, true);
    advance_block(p);
    RTB_CUDA(cudaGetLastError());
    return RTB_OK;
}

// A run of consecutive blocks in one call (offline rendering of a recorded session, benchmarks):
// block i reads d_in_ring + (i % n_ring) * S*C*B and d_yaw_ring + (i % n_ring) * S, every block
// writes d_out.  Same kernels and state updates as `count` calls of rtb_sh_process_block.
int rtb_sh_process_blocks(rtb_sh_plan *p, const float *d_in_ring, const float *d_yaw_ring, float *d_out,
                          int n_ring, long long first, int count, void *cuda_stream) {
    if (!p || !d_in_ring || !d_out || n_ring < 1 || count < 0 || first < 0)
        return fail(RTB_ERR_INVALID, "bad argument");
    const size_t in_blk = (size_t)p->S * p->C * p->B;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    auto run_direct = [&]() -> int {
        for (long long i = first; i < first + count; ++i) {
            const int r = (int)(i % n_ring);
            const int rc = rtb_sh_process_block(p, d_in_ring + (size_t)r * in_blk,
                                                d_yaw_ring ? d_yaw_ring + (size_t)r * p->S : nullptr, d_out, cuda_stream);
            if (rc != RTB_OK) return rc;
        }
        return RTB_OK;
    };
    // A run that comes back with the same buffers, ring phase, length and plan state is replayed as ONE
    // CUDA graph launch from its third occurrence on (built on the second): the device then walks the
    // run without the host in between -- with 8 ranks in one box the slowest rank of a 20-block run
    // was 4.5 % behind a single rank although every kernel took the same time.  Plain P = 1 plans on
    // an explicit stream only: nothing in their launch path depends on host state beyond what the key
    // holds, and advance_block() is all the host bookkeeping of a block.
    const bool eligible = p->graphs_enabled && count >= 4 && st != nullptr && st != cudaStreamLegacy &&
                          st != cudaStreamPerThread && p->P == 1 && p->tables_set && !p->profiling &&
                          (!p->mg_enabled || p->mg_order == p->cur_order) &&
                          (d_yaw_ring || p->passthrough) && p->last_valid == (p->crossfade && !p->passthrough);
    if (!eligible) return run_direct();
    DeviceGuard guard(p->device);
    rtb_sh_plan::BlockGraph key{};
    key.d_in = d_in_ring; key.d_yaw = d_yaw_ring; key.d_out = d_out; key.n_ring = n_ring;
    key.phase = (int)(first % n_ring); key.count = count; key.st = st;
    key.slot = p->slot; key.rad_idx = p->rad_idx; key.last_valid = p->last_valid; key.cur_order = p->cur_order;
    key.last_order = p->last_order; key.crossfade = p->crossfade; key.passthrough = p->passthrough;
    key.hp_slot = p->hp_P > 0 ? p->hp_slot : 0; key.hp_xidx = p->hp_P > 0 ? p->hp_xidx : 0; key.mg_order = p->mg_order;
    rtb_sh_plan::BlockGraph *g = nullptr;
    for (auto &e : p->graphs)
        if (e.d_in == key.d_in && e.d_yaw == key.d_yaw && e.d_out == key.d_out && e.n_ring == key.n_ring &&
            e.phase == key.phase && e.count == key.count && e.st == key.st && e.slot == key.slot &&
            e.rad_idx == key.rad_idx && e.last_valid == key.last_valid && e.cur_order == key.cur_order &&
            e.last_order == key.last_order && e.crossfade == key.crossfade && e.passthrough == key.passthrough &&
            e.hp_slot == key.hp_slot && e.hp_xidx == key.hp_xidx && e.mg_order == key.mg_order) { g = &e; break; }
    if (!g) {  // first occurrence: remember the key, launch directly
        if (p->graphs.size() >= 16) {  // a caller that never repeats itself: stop collecting
            for (auto &e : p->graphs) if (e.exec) cudaGraphExecDestroy(e.exec);
            p->graphs.clear();
        }
        key.seen = 1; key.exec = nullptr; key.launches = 0;
        p->graphs.push_back(key);
        return run_direct();
    }
    if (!g->exec) {  // second occurrence: record the run (the recording does not execute), then launch it
        // host state the recording advances; put back if the recording fails, and the run is launched
        // kernel by kernel instead (graphs off for this plan from then on)
        const int s_slot = p->slot, s_rad = p->rad_idx, s_lo = p->last_order, s_hps = p->hp_slot, s_hpx = p->hp_xidx;
        const bool s_lv = p->last_valid;
        const long long l0 = p->launches;
        cudaGraph_t graph = nullptr;
        cudaError_t ce = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
        int rc = RTB_OK;
        if (ce == cudaSuccess) {
            rc = run_direct();
            ce = cudaStreamEndCapture(st, &graph);
        }
        cudaError_t ie = cudaErrorUnknown;
        if (rc == RTB_OK && ce == cudaSuccess && graph) ie = cudaGraphInstantiate(&g->exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (ie != cudaSuccess) {
            (void)cudaGetLastError();
            g->exec = nullptr;
            p->graphs_enabled = false;
            p->slot = s_slot; p->rad_idx = s_rad; p->last_order = s_lo; p->hp_slot = s_hps; p->hp_xidx = s_hpx;
            p->last_valid = s_lv; p->launches = l0;
            return run_direct();
        }
        g->launches = p->launches - l0;
        g->seen = 2;
        RTB_CUDA(cudaGraphLaunch(g->exec, st));  // host state already stands where the run ends
        return RTB_OK;
    }
    RTB_CUDA(cudaGraphLaunch(g->exec, st));
    for (int i = 0; i < count; ++i) advance_block(p);
    p->launches += g->launches;
    p->graph_replays++;
    g->seen++;
    return RTB_OK;
}

namespace {

int host_slots_init(rtb_sh_plan *p) {
    if (p->hslots_ready) return RTB_OK;
    const size_t in_per = (size_t)p->C * p->B, out_per = (size_t)p->E * p->B;
    cudaError_t err = cudaSuccess;
    for (auto &h : p->hslot) {
        if (err == cudaSuccess) err = cudaMalloc(&h.d_in, sizeof(float) * p->S * in_per);
        if (err == cudaSuccess) err = cudaMalloc(&h.d_out, sizeof(float) * p->S * out_per);
        if (err == cudaSuccess) err = cudaMalloc(&h.d_yaw, sizeof(float) * p->S);
        if (err == cudaSuccess) err = cudaMemset(h.d_yaw, 0, sizeof(float) * p->S);
        if (err == cudaSuccess) err = cudaEventCreateWithFlags(&h.in_done, cudaEventDisableTiming);
        if (err == cudaSuccess) err = cudaEventCreateWithFlags(&h.k_done, cudaEventDisableTiming);
        if (err == cudaSuccess) err = cudaEventCreateWithFlags(&h.out_done, cudaEventDisableTiming);
    }
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&p->d2h_stream, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&p->k_stream, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaDeviceSynchronize();
    if (err != cudaSuccess) return fail(RTB_ERR_NOMEM, "staging allocation failed: %s", cudaGetErrorString(err));
    p->hslots_ready = true;
    return RTB_OK;
}

}  // namespace

// Host-buffer entry points.  A block travels through three streams -- input copy, kernels, result
// copy -- and RTB_HOST_SLOTS staging slots, so that consecutive submitted blocks form a software
// pipeline over the full-duplex link: while block t renders and its ear signals go back, the
// input of block t+1 is already on its way (the input copy is ~95 % of a step at PCIe Gen5 rates).
// The kernels of consecutive blocks stay ordered on one stream (they share the per-stream state).
int rtb_sh_submit_block_host(rtb_sh_plan *p, const float *h_in, const float *h_yaw_deg, float *h_out,
                             long long *ticket) {
    if (!p || !h_in || !h_out) return fail(RTB_ERR_INVALID, "NULL argument");
    if (!p->tables_set)
        return fail(RTB_ERR_STATE, "blocks in spherical harmonics domain have not been calculated yet.");
    if (!h_yaw_deg && !p->passthrough) return fail(RTB_ERR_INVALID, "NULL yaw pointer");
    DeviceGuard guard(p->device);
    if (int rc = host_slots_init(p)) return rc;
    const size_t in_per = (size_t)p->C * p->B, out_per = (size_t)p->E * p->B;
    rtb_sh_plan::HostSlot &h = p->hslot[p->submitted % RTB_HOST_SLOTS];
    const bool reused = p->submitted >= RTB_HOST_SLOTS;
    cudaStream_t st_in = p->own_stream, st_k = p->k_stream, st_out = p->d2h_stream;
    if (reused) {
        // the slot's previous block: its kernels have read d_in, its result copy has read d_out
        RTB_CUDA(cudaStreamWaitEvent(st_in, h.k_done, 0));
        RTB_CUDA(cudaStreamWaitEvent(st_k, h.out_done, 0));
    }
    if (h_yaw_deg)
        RTB_CUDA(cudaMemcpyAsync(h.d_yaw, h_yaw_deg, sizeof(float) * p->S, cudaMemcpyHostToDevice, st_in));
    RTB_CUDA(cudaMemcpyAsync(h.d_in, h_in, sizeof(float) * p->S * in_per, cudaMemcpyHostToDevice, st_in));
    RTB_CUDA(cudaEventRecord(h.in_done, st_in));
    RTB_CUDA(cudaStreamWaitEvent(st_k, h.in_done, 0));
    launch_range(p, 0, p->S, h.d_in, h.d_yaw, h.d_out, st_k);
    RTB_CUDA(cudaEventRecord(h.k_done, st_k));
    RTB_CUDA(cudaStreamWaitEvent(st_out, h.k_done, 0));
    RTB_CUDA(cudaMemcpyAsync(h_out, h.d_out, sizeof(float) * p->S * out_per, cudaMemcpyDeviceToHost, st_out));
    RTB_CUDA(cudaEventRecord(h.out_done, st_out));
    advance_block(p);
    if (ticket) *ticket = p->submitted;
    p->submitted++;
    RTB_CUDA(cudaGetLastError());
    return RTB_OK;
}

int rtb_sh_wait_block_host(rtb_sh_plan *p, long long ticket) {
    if (!p) return fail(RTB_ERR_INVALID, "NULL plan");
    if (ticket < 0 || ticket >= p->submitted) return fail(RTB_ERR_INVALID, "unknown ticket %lld", ticket);
    // (a slot that has been reused since carries a later block's event: the result stream is in
    // order, so waiting for it covers this block too)
    DeviceGuard guard(p->device);
    RTB_CUDA(cudaEventSynchronize(p->hslot[ticket % RTB_HOST_SLOTS].out_done));
    return RTB_OK;
}

int rtb_sh_process_block_host(rtb_sh_plan *p, const float *h_in, const float *h_yaw_deg, float *h_out) {
    long long ticket = 0;
    if (int rc = rtb_sh_submit_block_host(p, h_in, h_yaw_deg, h_out, &ticket)) return rc;
    return rtb_sh_wait_block_host(p, ticket);
}

int rtb_sh_plan_query(rtb_sh_plan *p, int what, int64_t *value) {
    if (!p || !value) return fail(RTB_ERR_INVALID, "NULL argument");
    switch (what) {
        case RTB_Q_SYMMETRIC: *value = p->symmetric ? 1 : 0; break;
        case RTB_Q_ROWS: *value = p->J; break;
        case RTB_Q_SIGNALS: *value = p->NP; break;
        case RTB_Q_LAUNCHES: *value = p->launches; break;
        case RTB_Q_TENSOR_CORE: *value = p->tc_enabled ? 1 : 0; break;
        case RTB_Q_RENDER_KERNEL: *value = p->last_render_kernel; break;
        case RTB_Q_GRAPH_REPLAYS: *value = p->graph_replays; break;
        case RTB_Q_STATE_BYTES: *value = (int64_t)sizeof(float) * (2 * (int64_t)p->J * p->B + 2); break;
        case RTB_Q_ANALYSIS_NS:
        case RTB_Q_RENDER_NS: {
            if (!p->profiling || !p->ev_valid) return fail(RTB_ERR_STATE, "profiling is off or no block was processed");
            DeviceGuard guard(p->device);
            RTB_CUDA(cudaEventSynchronize(p->ev[2]));
            float ms = 0.f;
            const int i = what == RTB_Q_ANALYSIS_NS ? 0 : 1;
            RTB_CUDA(cudaEventElapsedTime(&ms, p->ev[i], p->ev[i + 1]));
            *value = (int64_t)(ms * 1e6);
            break;
        }
        default: return fail(RTB_ERR_INVALID, "unknown query %d", what);
    }
    return RTB_OK;
}

// Set-up time: SH transform of an HRIR set, H_nm[p][k][e][f] = sum_c Y_w[k][c] * H[p][c][e][f]
// (FilterSetMiro.calculate_filter_blocks_nm, _filter_set.py:1093-1149: a [K x directions] .
// [directions x F] complex product per partition and ear; 2702 directions for the reference's
// default HRIR set).  Runs on the tcgen05 analysis kernel: the real and imaginary parts of Y_w are
// the analysis rows, the directions take the role of the array channels, the bins of (partition,
// ear, re | im) the role of a stream's samples; TF32 x 3 split as everywhere (2e-6 relative).
// The Nyquist bin (block length + 1 bins, the kernel works on powers of two) is summed on the host.
int rtb_sh_transform_filter(int device, const float *yw, const float *hfd, int n_partitions, int n_coeffs,
                            int n_dirs, int n_ears, int n_bins, float *hnm) {
    if (!yw || !hfd || !hnm) return fail(RTB_ERR_INVALID, "NULL argument");
    const int P = n_partitions, K = n_coeffs, C = n_dirs, E = n_ears, F = n_bins, Bq = n_bins - 1;
    if (P < 1 || K < 1 || C < 1 || E < 1) return fail(RTB_ERR_INVALID, "sizes must be >= 1");
    if (Bq < 32 || (Bq & (Bq - 1))) return fail(RTB_ERR_UNSUPPORTED, "n_bins - 1 must be a power of two >= 32 (got %d)", Bq);
    int ndev = 0;
    RTB_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(RTB_ERR_INVALID, "no CUDA device %d", device);
    DeviceGuard guard(device);
    const int S = P * E * 2;  // "streams": (partition, ear, re | im)
    const int KC = RTB_TC_KC;
    // The tensor core adds into its fp32 accumulator with truncation, so the error of one long sum
    // grows with the number of k-steps (measured 2e-5 relative over 2702 directions in one pass).
    // The directions are therefore summed in chunks of 256 on the device and the partial sums are
    // added in double precision on the host (1e-6 relative).
    const int CH = 256;
    const int rows_total = 2 * K, rows_per = 256;
    std::vector<double> zall((size_t)S * rows_total * Bq, 0.0);
    std::vector<float> x((size_t)S * CH * Bq), z((size_t)S * std::min(rows_total, rows_per) * Bq);
    float *d_x = nullptr, *d_z = nullptr, *d_b = nullptr;
    cudaError_t err = cudaMalloc(&d_x, sizeof(float) * x.size());
    if (err == cudaSuccess) err = cudaMalloc(&d_z, sizeof(float) * z.size());
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
    for (int c0 = 0; c0 < C && err == cudaSuccess; c0 += CH) {
        const int Cc = std::min(CH, C - c0), Kpad = (Cc + KC - 1) / KC * KC;
        for (int p = 0; p < P; ++p)        // input chunk [S][Cc][Bq]
            for (int c = 0; c < Cc; ++c)
                for (int e = 0; e < E; ++e) {
                    const float *src = hfd + 2 * ((((size_t)p * C + c0 + c) * E + e) * F);
                    float *dr = x.data() + (((size_t)(p * E + e) * 2 + 0) * Cc + c) * Bq;
                    float *di = x.data() + (((size_t)(p * E + e) * 2 + 1) * Cc + c) * Bq;
                    for (int f = 0; f < Bq; ++f) { dr[f] = src[2 * f]; di[f] = src[2 * f + 1]; }
                }
        err = cudaMemcpy(d_x, x.data(), sizeof(float) * (size_t)S * Cc * Bq, cudaMemcpyHostToDevice);
        for (int r0 = 0; r0 < rows_total && err == cudaSuccess; r0 += rows_per) {
            const int J = std::min(rows_per, rows_total - r0), Npad = (J + 15) / 16 * 16;
            std::vector<float> bsrc(2 * (size_t)Kpad * Npad, 0.f);
            for (int j = 0; j < J; ++j) {
                const int row = r0 + j, k = row < K ? row : row - K, ri = row < K ? 0 : 1;
                for (int c = 0; c < Cc; ++c) {
                    const float v = yw[2 * ((size_t)k * C + c0 + c) + ri];
                    uint32_t u;
                    memcpy(&u, &v, 4);
                    u = (u + 0x1000u) & 0xffffe000u;  // round to nearest TF32
                    float hi;
                    memcpy(&hi, &u, 4);
                    const size_t off = (((size_t)(c / 4) * (Npad / 8) + j / 8) * 8 + j % 8) * 4 + c % 4;
                    bsrc[off] = hi;
                    bsrc[(size_t)Kpad * Npad + off] = v - hi;
                }
            }
            cudaFree(d_b);
            d_b = nullptr;
            err = cudaMalloc(&d_b, sizeof(float) * bsrc.size());
            if (err == cudaSuccess) err = cudaMemcpy(d_b, bsrc.data(), sizeof(float) * bsrc.size(), cudaMemcpyHostToDevice);
            if (err != cudaSuccess) break;
            const size_t a_stage = sizeof(float) * 2 * KC * RTB_TC_M, b_stage = sizeof(float) * 2 * (size_t)KC * Npad;
            const int nst = (216 * 1024) / (a_stage + b_stage) >= 4 ? 4 : 2;
            const size_t smem = nst * (a_stage + b_stage);
            const int cols = Npad <= 32 ? 32 : (Npad <= 64 ? 64 : (Npad <= 128 ? 128 : 256));
            const long long tiles = ((long long)S * Bq + RTB_TC_M - 1) / RTB_TC_M;
            const int grid = (int)std::min<long long>(tiles, n_sm);
            kernel_attrs_once(sh_analysis_tc_kernel<true, RTB_TC_LG>, 224 * 1024);
            sh_analysis_tc_kernel<true, RTB_TC_LG><<<grid, RTB_TC_THREADS_OF(RTB_TC_LG), smem, 0>>>(
                d_x, d_b, d_z, S, Cc, J, Bq, Kpad, Npad, cols, nst, 0LL);
            err = cudaGetLastError();
            if (err == cudaSuccess) err = cudaMemcpy(z.data(), d_z, sizeof(float) * (size_t)S * J * Bq, cudaMemcpyDeviceToHost);
            if (err != cudaSuccess) break;
            for (int s = 0; s < S; ++s) {
                double *dst = zall.data() + ((size_t)s * rows_total + r0) * Bq;
                const float *srcz = z.data() + (size_t)s * J * Bq;
                for (size_t i = 0; i < (size_t)J * Bq; ++i) dst[i] += srcz[i];
            }
        }
    }
    cudaFree(d_x); cudaFree(d_z); cudaFree(d_b);
    if (err != cudaSuccess) return fail(RTB_ERR_CUDA, "SH transform of the filter set failed: %s", cudaGetErrorString(err));
    for (int p = 0; p < P; ++p)
        for (int k = 0; k < K; ++k)
            for (int e = 0; e < E; ++e) {
                float *dst = hnm + 2 * ((((size_t)p * K + k) * E + e) * F);
                const double *zr = zall.data() + ((size_t)((p * E + e) * 2 + 0) * rows_total) * Bq;  // input = Re H
                const double *zi = zall.data() + ((size_t)((p * E + e) * 2 + 1) * rows_total) * Bq;  // input = Im H
                const double *yr_hr = zr + (size_t)k * Bq, *yi_hr = zr + (size_t)(K + k) * Bq;
                const double *yr_hi = zi + (size_t)k * Bq, *yi_hi = zi + (size_t)(K + k) * Bq;
                for (int f = 0; f < Bq; ++f) {
                    dst[2 * f] = (float)(yr_hr[f] - yi_hi[f]);
                    dst[2 * f + 1] = (float)(yr_hi[f] + yi_hr[f]);
                }
                double ar = 0.0, ai = 0.0;  // Nyquist bin
                for (int c = 0; c < C; ++c) {
                    const double yr = yw[2 * ((size_t)k * C + c)], yi = yw[2 * ((size_t)k * C + c) + 1];
                    const float *h = hfd + 2 * ((((size_t)p * C + c) * E + e) * F + Bq);
                    ar += yr * h[0] - yi * h[1];
                    ai += yr * h[1] + yi * h[0];
                }
                dst[2 * Bq] = (float)ar;
                dst[2 * Bq + 1] = (float)ai;
            }
    return RTB_OK;
}

#ifdef RTB_TRACE_TC
int rtb_debug_tc_trace(long long *h_out, int n) {  // trace builds only (not part of the public ABI)
    return cudaMemcpyFromSymbol(h_out, rtb::g_tc_trace, sizeof(long long) * n) == cudaSuccess ? RTB_OK : RTB_ERR_CUDA;
}
#endif
int rtb_sh_plan_destroy(rtb_sh_plan *p) {
    if (p) drop_graphs(p);
    if (!p) return RTB_OK;
    DeviceGuard guard(p->device);
    free_tables(p);
    cudaFree(p->d_tw); cudaFree(p->d_win_in); cudaFree(p->d_win_out); cudaFree(p->d_radlast);
    cudaFree(p->d_hp_hfd); cudaFree(p->d_hp_hist); cudaFree(p->d_hp_x);
    for (auto &h : p->hslot) {
        cudaFree(h.d_in); cudaFree(h.d_yaw); cudaFree(h.d_out);
        if (h.in_done) cudaEventDestroy(h.in_done);
        if (h.k_done) cudaEventDestroy(h.k_done);
        if (h.out_done) cudaEventDestroy(h.out_done);
    }
    if (p->own_stream) cudaStreamDestroy(p->own_stream);
    if (p->d2h_stream) cudaStreamDestroy(p->d2h_stream);
    if (p->k_stream) cudaStreamDestroy(p->k_stream);
    for (auto &e : p->ev) if (e) cudaEventDestroy(e);
    delete p;
    return RTB_OK;
}

}  // extern "C"
