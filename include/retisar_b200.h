/*
 * retisar_b200 -- C ABI of the B200-native per-block binaural rendering chain.
 *
 * The reference (AppliedAcousticsChalmers/ReTiSAR, pure Python) has no FFI; its boundary for this
 * path is the duck-typed class API of ReTiSAR/_convolver.py.  Every entry point below names the
 * reference interface it replaces.  Plain pointers and sizes only; the library links against the
 * CUDA runtime and nothing else (no torch types).
 *
 * Conventions
 *   - complex arrays are interleaved float pairs (re, im) == numpy complex64, C order
 *   - "d_" pointers are device pointers on the plan's device, "h_" pointers are host pointers
 *   - every function returns RTB_OK or an RTB_ERR_* code; rtb_last_error() gives the message of
 *     the last failure on the calling thread.  The Python layer maps codes to the exceptions the
 *     reference raises (ValueError / RuntimeError / NotImplementedError).
 *   - a plan batches S independent render streams (reference: one convolver object per stream,
 *     ReTiSAR/_jack_renderer.py:467-531 for the benchmark list of convolvers)
 */
#ifndef RETISAR_B200_H
#define RETISAR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTB_OK 0
#define RTB_ERR_INVALID 1     /* bad argument            -> ValueError          */
#define RTB_ERR_STATE 2       /* tables not set etc.     -> RuntimeError        */
#define RTB_ERR_UNSUPPORTED 3 /* not implemented         -> NotImplementedError */
#define RTB_ERR_CUDA 4        /* CUDA runtime failure    -> RuntimeError        */
#define RTB_ERR_NOMEM 5       /* allocation failure      -> MemoryError         */

typedef struct rtb_sh_plan rtb_sh_plan;   /* AdjustableShConvolver x S streams */
typedef struct rtb_ols_plan rtb_ols_plan; /* OverlapSaveConvolver  x S streams */

/* library ------------------------------------------------------------------------------------ */
int rtb_version(void);
const char *rtb_last_error(void);
/* number of visible CUDA devices, or a negative RTB error code */
int rtb_device_count(void);

/* SH-domain renderer ------------------------------------------------------------------------- */
/* AdjustableShConvolver.__init__ (_convolver.py:1010-1049) for S batched streams.
 * block_length B: a power of two, 32 ... 4096 (32 ... 256 when n_partitions > 1); n_ears must be 2; n_partitions = number of B-sample
 * partitions of the HRIR.  1 is the reference's only supported case (_filter_set.py:1136-1140);
 * > 1 is an EXTENSION defined by composition with the reference's partition-queue semantics
 * (_convolver.py:802-836, :645-665), see DESIGN.md row a-X. */
int rtb_sh_plan_create(rtb_sh_plan **plan, int device, int n_streams, int block_length,
                       int n_channels, int sh_max_order, int n_ears, int n_partitions);

/* prepare_sh_processing (_convolver.py:1071-1120): upload the tables the hot path reads.
 *   yw     [K][C]       complex  _sh_bases_weighted            (_convolver.py:1097)
 *   hnm    [P][K][E][F] complex  _filter._irs_blocks_nm, already multiplied with the
 *                                compensation spectra            (_convolver.py:1206)
 *   sh_m   [K]          int16    _sh_m                         (_convolver.py:1094)
 *   rev    [K]          int32    _sh_m_rev_id                  (_convolver.py:1095)
 *   win_in / win_out [B] float   cross-fade windows (_convolver.py:741-747); NULL = cos^2 default
 * Resets all per-stream state like the reference (input buffer :1113-1116, last phases :1098). */
int rtb_sh_plan_set_tables(rtb_sh_plan *plan, const float *yw, const float *hnm,
                           const int16_t *sh_m, const int32_t *rev, const float *win_in,
                           const float *win_out);

/* _apply_sh_compensation re-run by update_sh_processing (_convolver.py:1161-1162, :1189-1206):
 * replace hnm [P][K][E][F] only; per-stream state (overlap buffer, last yaw) is kept. */
int rtb_sh_plan_update_filter(rtb_sh_plan *plan, const float *hnm);

/* _sources_deg[0,0] (_convolver.py:729) and tracker_dir = +1 HRIR / -1 BRIR (_convolver.py:958) */
int rtb_sh_plan_set_source(rtb_sh_plan *plan, float source_azim_deg, int tracker_dir);
/* update_sh_processing: _sh_cur_order (_convolver.py:1159-1160); 0 <= order <= sh_max_order */
int rtb_sh_plan_set_order(rtb_sh_plan *plan, int sh_cur_order);
/* Convolver.set_passthrough (_convolver.py:288-305) */
int rtb_sh_plan_set_passthrough(rtb_sh_plan *plan, int on);
/* AdjustableShConvolver.set_crossfade (_convolver.py:1318-1341); off clears the "last" state */
int rtb_sh_plan_set_crossfade(rtb_sh_plan *plan, int on);
/* _clear_buffers (_convolver.py:774-778) */
int rtb_sh_plan_reset(rtb_sh_plan *plan);

/* AdjustableShConvolver.filter_block (_convolver.py:1231-1316) for all S streams.
 *   d_in  [S][C][B] float32, d_yaw_deg [S] float32 (tracker AZIM in degrees, _tracker.py:73-81),
 *   d_out [S][E][B] float32; d_in and d_out 16-byte aligned (RTB_ERR_INVALID otherwise).
 *   Asynchronous on `cuda_stream` (a cudaStream_t, NULL = default). */
int rtb_sh_process_block(rtb_sh_plan *plan, const float *d_in, const float *d_yaw_deg,
                         float *d_out, void *cuda_stream);
/* `count` consecutive blocks in one call: block i (first <= i < first + count) reads
 * d_in_ring + (i % n_ring) * S*C*B and d_yaw_ring + (i % n_ring) * S; every block writes d_out. */
int rtb_sh_process_blocks(rtb_sh_plan *plan, const float *d_in_ring, const float *d_yaw_ring,
                          float *d_out, int n_ring, long long first, int count, void *cuda_stream);
/* Same with host buffers: H2D copy, render, D2H copy, synchronise (the JACK-callback shape of
 * JackRenderer._process, _jack_renderer.py:296-315). */
int rtb_sh_process_block_host(rtb_sh_plan *plan, const float *h_in, const float *h_yaw_deg,
                              float *h_out);
/* The same block split into an asynchronous submit and a wait, for callers that feed blocks
 * back to back (the reference's real-time feeder hands a block to JackRenderer._process every
 * period, _jack_client.py:713-867): up to three submitted blocks are in flight, the input copy of
 * block t+1 overlapping the kernels and the result copy of block t.  h_in / h_yaw_deg must stay
 * valid until the block's wait returns (page-locked memory, rtb_host_alloc, makes the copies
 * direct DMA); h_out is complete when rtb_sh_wait_block_host(plan, *ticket) returns.  Blocks are
 * processed in submission order. */
int rtb_sh_submit_block_host(rtb_sh_plan *plan, const float *h_in, const float *h_yaw_deg,
                             float *h_out, long long *ticket);
int rtb_sh_wait_block_host(rtb_sh_plan *plan, long long ticket);

/* Pre-rendered mode (extension).  The array signals enter the chain only through the analysis rows
 * z[j][t] = sum_c R[j][c] x[c][t] (R = real / imaginary parts of _sh_bases_weighted,
 * _convolver.py:1260-1263 applied before the FFT).  When x is itself the output of the reference's
 * ARIR pre-renderer -- a mono source convolved with one impulse response per array channel
 * (_run.py:240-288) -- R folds into those filters at set-up time and the pre-renderer can write the J
 * rows straight into the plan:
 *   rtb_sh_plan_get_analysis_rows  rows [J][C] float, J = RTB_Q_ROWS
 *   rtb_sh_plan_rows_buffer        device buffer [S][J][B] that must hold the rows of the NEXT block
 *   rtb_sh_process_block_rows      renders that block without the analysis GEMM */
int rtb_sh_plan_get_analysis_rows(rtb_sh_plan *plan, float *rows);
int rtb_sh_plan_rows_buffer(rtb_sh_plan *plan, float **d_rows);
int rtb_sh_process_block_rows(rtb_sh_plan *plan, const float *d_yaw_deg, float *d_out, void *cuda_stream);

/* Optional headphone-compensation filter (HPCF) behind the renderer.  The reference chains it as a
 * separate client after the renderer (_run.py:290-338): a 2-in / 2-out channel-wise
 * OverlapSaveConvolver (_convolver.py:524-571) fed with the binaural block.  Here the warp that
 * produced a stream's ear signals continues with the filter in the same launch (input-side delay
 * line of n_partitions binaural spectra per stream).  hfd [n_partitions][2][F] complex =
 * FilterSet.get_filter_blocks_fd() of the HPCF set (_filter_set.py:625-664); n_partitions == 0
 * (hfd may be NULL) detaches the filter.  Attaching resets the filter's own state only. */
int rtb_sh_plan_set_hpcf(rtb_sh_plan *plan, const float *hfd, int n_partitions);

/* Set-up time: FilterSetMiro.calculate_filter_blocks_nm (_filter_set.py:1093-1149) on the device,
 * H_nm[p][k][e][f] = sum_c Y_w[k][c] * H[p][c][e][f] -- the SH transform of an HRIR set
 * ([K x directions] . [directions x F] per partition and ear) on the tensor-core analysis kernel.
 * Host pointers: yw [K][C] complex, hfd [P][C][E][F] complex (get_filter_blocks_fd layout),
 * hnm [P][K][E][F] complex (out); F - 1 must be a power of two >= 32.  TF32 x 3: ~2e-6 relative
 * to the numpy product of the reference. */
int rtb_sh_transform_filter(int device, const float *yw, const float *hfd, int n_partitions,
                            int n_coeffs, int n_dirs, int n_ears, int n_bins, float *hnm);

/* introspection: what = RTB_Q_* */
#define RTB_Q_SYMMETRIC 1   /* 1 if Y_w has the (n,-m) conjugate symmetry (real analysis path) */
#define RTB_Q_ROWS 2        /* J: real analysis rows per stream                                */
#define RTB_Q_SIGNALS 3     /* number of complex FFT signals per stream                        */
#define RTB_Q_LAUNCHES 4    /* kernels launched by this plan so far                            */
#define RTB_Q_STATE_BYTES 5 /* bytes of per-stream device state                                */
#define RTB_Q_ANALYSIS_NS 6 /* device time of the last block's analysis kernel (profiling on)  */
#define RTB_Q_RENDER_NS 7   /* device time of the last block's render kernel   (profiling on)  */
#define RTB_Q_TENSOR_CORE 8 /* 1 if the analysis GEMM runs on tcgen05 (wide arrays)             */
#define RTB_Q_RENDER_KERNEL 9 /* render kernel of the last block: 0 sh_render_kernel (term-wise),
                                 1 sh_render_mg_kernel (grouped by m),
                                 3 partitioned (spectra + SIMT partition MAC + tail), 4 passthrough,
                                 6 partitioned on tensor cores (sh_partition_tc_kernel)            */
#define RTB_Q_GRAPH_REPLAYS 10 /* runs of blocks replayed as one CUDA graph launch so far
                                  (rtb_sh_process_blocks)                                        */
int rtb_sh_plan_query(rtb_sh_plan *plan, int what, int64_t *value);
/* record CUDA events around each kernel of process_block (measurement aid, no reference
 * counterpart; the reference reports load through JACK, _jack_client.py:869-895) */
int rtb_sh_plan_set_profiling(rtb_sh_plan *plan, int on);
int rtb_sh_plan_destroy(rtb_sh_plan *plan);

/* Uniformly partitioned overlap-save FIR convolver ------------------------------------------- */
/* OverlapSaveConvolver.__init__ (_convolver.py:454-480): channel-wise FIR, n_in == n_out or
 * n_in == 1 (mono input broadcast to all filter channels, _convolver.py:588-591). */
int rtb_ols_plan_create(rtb_ols_plan **plan, int device, int n_streams, int block_length,
                        int n_in, int n_out, int n_partitions);
/* filter spectra _irs_blocks_fd [P][n_out][F] complex (FilterSet.calculate_filter_blocks_fd,
 * _filter_set.py:625-664; the singleton direction axis dropped) */
int rtb_ols_plan_set_filter(rtb_ols_plan *plan, const float *hfd);
int rtb_ols_plan_set_passthrough(rtb_ols_plan *plan, int on);
int rtb_ols_plan_reset(rtb_ols_plan *plan);
/* OverlapSaveConvolver.filter_block (_convolver.py:524-571): d_in [S][n_in][B] -> d_out [S][n_out][B] */
int rtb_ols_process_block(rtb_ols_plan *plan, const float *d_in, float *d_out, void *cuda_stream);
int rtb_ols_process_block_host(rtb_ols_plan *plan, const float *h_in, float *h_out);
int rtb_ols_plan_query(rtb_ols_plan *plan, int what, int64_t *value); /* RTB_Q_LAUNCHES, _STATE_BYTES, _RENDER_NS */
int rtb_ols_plan_set_profiling(rtb_ols_plan *plan, int on);
int rtb_ols_plan_destroy(rtb_ols_plan *plan);

/* HRIR / BRIR switching renderer ----------------------------------------------------------------- */
/* AdjustableFdConvolver.__init__ (_convolver.py:702-751) for S batched streams: n_sources input
 * channels per stream, a filter set of n_dirs horizontal directions (FilterSetSsr: 360,
 * _filter_set.py:843-899) with n_partitions blocks, 2 ears. */
typedef struct rtb_fd_plan rtb_fd_plan;
int rtb_fd_plan_create(rtb_fd_plan **plan, int device, int n_streams, int block_length, int n_sources,
                       int n_dirs, int n_partitions);
/* _irs_blocks_fd [P][n_dirs][2][F] complex (_filter_set.py:625-664) and the cross-fade windows
 * (_convolver.py:741-747; NULL = cos^2 default) */
int rtb_fd_plan_set_filter(rtb_fd_plan *plan, const float *hfd, const float *win_in, const float *win_out);
/* _sources_deg[:, 0] (_convolver.py:729) and tracker_dir = +1 HRIR / -1 BRIR (_convolver.py:958) */
int rtb_fd_plan_set_sources(rtb_fd_plan *plan, const float *source_azim_deg, int tracker_dir);
int rtb_fd_plan_set_passthrough(rtb_fd_plan *plan, int on);
/* set_crossfade (_convolver.py:868-892): off clears the previous-filter queue */
int rtb_fd_plan_set_crossfade(rtb_fd_plan *plan, int on);
int rtb_fd_plan_reset(rtb_fd_plan *plan);
/* AdjustableFdConvolver.filter_block (_convolver.py:780-836): d_in [S][n_sources][B], d_yaw_deg [S]
 * -> d_out [S][2][B] */
int rtb_fd_process_block(rtb_fd_plan *plan, const float *d_in, const float *d_yaw_deg, float *d_out,
                         void *cuda_stream);
int rtb_fd_process_block_host(rtb_fd_plan *plan, const float *h_in, const float *h_yaw_deg, float *h_out);
int rtb_fd_plan_query(rtb_fd_plan *plan, int what, int64_t *value);
int rtb_fd_plan_destroy(rtb_fd_plan *plan);

/* ---- page-locked host buffers for the *_process_block_host entry points (the reference hands
 * numpy blocks to filter_block, ReTiSAR/_jack_renderer.py:315; pinned blocks make the host-to-device
 * copy a single DMA).  write_combined != 0: write-combined memory, for input blocks only. */
int rtb_host_alloc(void **ptr, unsigned long long bytes, int write_combined);
int rtb_host_free(void *ptr);

/* ---- output stage of a client (reference: JackClient._process_deliver,
 * ReTiSAR/_jack_client.py:770-867; levels as tools.calculate_rms / calculate_peak,
 * tools.py:1117-1169).  In place: d_out[s][c][:] *= d_gain[c] (output volume x per-channel
 * relative volume); d_rms / d_peak [n_streams][n_channels] receive the level of the scaled block
 * (dB with zeros reported as -200 when as_level != 0; either may be NULL); *d_clip_count is
 * incremented once per channel block whose peak exceeds 1 (the reference logs a warning);
 * d_clip_count may be NULL when clipping detection is off. */
int rtb_output_stage(float *d_out, const float *d_gain, float *d_rms, float *d_peak,
                     unsigned int *d_clip_count, long long n_streams, int n_channels,
                     int block_length, int as_level, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* RETISAR_B200_H */
