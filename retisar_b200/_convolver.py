"""Block-wise convolvers -- host-side mirror of ``ReTiSAR/_convolver.py`` over the CUDA library.

Class names, the factory, method signatures and error behaviour follow the reference so that
``JackRenderer`` (``_jack_renderer.py``) and ``_run_benchmark`` can use these classes unchanged;
the arithmetic of ``filter_block`` runs in ``libretisar_b200.so`` (no CPU path exists).

Extension over the reference: one convolver object may carry ``n_streams`` independent render
streams.  ``filter_block`` then takes ``[n_streams; channels; block]`` blocks (numpy on the host
or torch tensors on the device) and, for the binaural renderer, one yaw angle per stream.
"""
from copy import copy

import numpy as np

from . import config
from ._filter_set import FilterSetMiro, FilterSetMultiChannel, FilterSetSofa, FilterSetSsr  # noqa: F401
from ._tracker import HeadTracker
from .plans import FdPlan, OlsPlan, ShPlan
from . import sph


def _is_torch_tensor(x):
    return type(x).__module__.startswith("torch")


class DelayBuffer(object):
    """Delay line of time-domain blocks in a ring buffer (reference ``DelayBuffer``,
    ``_convolver.py:11-141``): the realised delay is a whole number of blocks, limited by the ring
    size given to ``init``.  Same attributes and return values as the reference, including its
    two quirks: a delay equal to the ring size returns the block just written (python index
    ``-len`` wraps to 0, ``:134``), and ``set_delay`` before ``init`` returns the delay in blocks.

    Extension: the blocks may carry a leading stream axis (``[n_streams; channels; block]``) and
    may be torch CUDA tensors; the ring then lives in device memory (``init(..., device=...)``)
    and ``process_block`` costs one device-to-device block copy on the current stream.
    """

    def __init__(self, sample_rate, block_length, delay_ms=0.0):
        self._buffer = None
        self._buffer_ptr = 0
        self._sample_rate = sample_rate
        self._block_length = block_length
        self._delay_blocks = self._calculate_delay_blocks(delay_ms)

    def init(self, max_length_sec, channel_count, n_streams=None, device=None):
        # +1 to realise the maximum delay without wrapping around to the current block (`:63-66`)
        block_count = self._calculate_delay_blocks(delay_ms=max_length_sec * 1000) + 1
        shape = (block_count,) + (() if n_streams is None else (n_streams,)) + (channel_count, self._block_length)
        if device is None:
            self._buffer = np.zeros(shape=shape, dtype=np.float32)
        else:
            import torch

            self._buffer = torch.zeros(shape, dtype=torch.float32, device=device)
        self._reset()

    def _calculate_delay_blocks(self, delay_ms):
        return int(np.ceil((delay_ms / 1000) * self._sample_rate / self._block_length))

    def _reset(self):
        if self._buffer is not None:
            if _is_torch_tensor(self._buffer):
                self._buffer.zero_()
            else:
                self._buffer.fill(0)
        self._buffer_ptr = 0

    def set_delay(self, new_delay_ms=0.0):
        if self._buffer is None:
            return self._delay_blocks
        self._delay_blocks = self._calculate_delay_blocks(new_delay_ms)  # round to full blocks
        if self._delay_blocks > self._buffer.shape[0]:
            self._delay_blocks = self._buffer.shape[0]
        if self._delay_blocks == 0:
            self._reset()
        return (self._delay_blocks * self._block_length / self._sample_rate) * 1000

    def process_block(self, input_td):
        if self._buffer is None or not self._delay_blocks or input_td is None:
            return input_td
        if _is_torch_tensor(self._buffer):
            self._buffer[self._buffer_ptr].copy_(input_td, non_blocking=True)
        else:
            self._buffer[self._buffer_ptr] = input_td
        # a view into the ring, like the reference (valid until the slot is written again)
        output_td = self._buffer[self._buffer_ptr - self._delay_blocks]
        self._buffer_ptr = (self._buffer_ptr + 1) % self._buffer.shape[0]
        return output_td


class Convolver(object):
    """Base class and factory (reference ``Convolver``, ``_convolver.py:144-433``)."""

    @staticmethod
    def create_instance_by_filter_set(filter_set, block_length=None, source_positions=None,
                                      shared_tracker_data=None, n_streams=1, device=None):
        """Pick the convolver class for a filter set (``_convolver.py:165-218``)."""
        if not block_length:
            convolver = Convolver(filter_set)
        elif type(filter_set) == FilterSetMultiChannel:
            convolver = OverlapSaveConvolver(filter_set, block_length, n_streams=n_streams,
                                             device=device)
        elif type(filter_set) == FilterSetSsr:
            convolver = AdjustableFdConvolver(filter_set, block_length, source_positions,
                                              shared_tracker_data, n_streams=n_streams,
                                              device=device)
        elif filter_set.get_sh_configuration().arir_config:
            convolver = OverlapSaveConvolver(filter_set, block_length, n_streams=n_streams,
                                             device=device)
        else:
            convolver = AdjustableShConvolver(filter_set, block_length, source_positions,
                                              shared_tracker_data, n_streams=n_streams,
                                              device=device)
        if config.IS_DEBUG_MODE and type(convolver) is not AdjustableShConvolver:
            convolver._debug_filter_block(len(source_positions) if source_positions else 0)
        return convolver

    def __init__(self, filter_set):
        self._filter = filter_set
        self._is_passthrough = False
        self._rfft = None  # kept for attribute compatibility; transforms run on the GPU
        self._irfft = None
        if type(self) is Convolver:
            self._filter.calculate_filter_blocks_fd(None)

    def __str__(self):
        return f"[ID={id(self)}, _filter={self._filter}, _is_passthrough={self._is_passthrough}]"

    def _clear_buffers(self):
        """Zero all signal state (``_convolver.py:257-259``)."""

    def init_fft_optimize(self, logger=None):
        """The reference plans its FFTW transforms here (``_convolver.py:261-286``); the CUDA
        plans are complete once the tables are uploaded, so this only reports."""
        msg = "skipping FFTW DFT optimization (transforms run in CUDA kernels)."
        logger.info(msg) if logger else print(msg)

    def set_passthrough(self, new_state=None):
        """Toggle / set bypass (``_convolver.py:288-305``)."""
        self._is_passthrough = (not self._is_passthrough) if new_state is None else bool(new_state)
        self._apply_passthrough()
        return self._is_passthrough

    def _apply_passthrough(self):
        pass

    def _debug_filter_block(self, input_count, is_generate_noise=False):
        """Push one dirac (or noise) block through ``filter_block`` (``_convolver.py:307-361``)."""
        if not input_count:
            input_count = self.get_input_channel_count()
        dtype = self._filter.get_dirac_td().dtype
        if is_generate_noise:
            ip = (np.random.default_rng(0).standard_normal((input_count, self._block_length)) * 0.1
                  ).astype(dtype)
        else:
            ip = np.zeros((input_count, self._block_length), dtype=dtype)
            ip[:, 0] = 1
        return self.filter_block(ip)

    def filter_block(self, input_td):
        """Un-buffered circular convolution of the reference base class (``_convolver.py:363-399``).
        It is documented there as "not recommended" and is not on the render path; only the
        pause protocol (``None`` in, ``None`` out) is kept."""
        if input_td is None:
            if not config.IS_RUNNING.is_set():
                self._clear_buffers()
            return None
        raise NotImplementedError(
            "the un-buffered base `Convolver` has no CUDA implementation; pass a block length")

    def get_input_channel_count(self):
        """``_convolver.py:402-412``"""
        shape = self._filter._irs_td.shape
        return shape[1] if self._filter._is_hpcf else shape[0]

    def get_output_channel_count(self):
        """``_convolver.py:414-422``"""
        return self._filter._irs_td.shape[0] * self._filter._irs_td.shape[1]

    def _get_current_filters_fd(self):
        if self._is_passthrough:
            return self._filter.get_dirac_blocks_fd()
        return self._filter.get_filter_blocks_fd()


class OverlapSaveConvolver(Convolver):
    """Uniformly partitioned overlap-save FIR convolver (``_convolver.py:436-669``).

    Channel-wise: output channel i is input channel i (or the single input channel, which the
    reference broadcasts, ``:588-591``) convolved with filter channel i.
    """

    def __init__(self, filter_set, block_length, n_streams=1, device=None):
        super().__init__(filter_set=filter_set)
        self._block_length = block_length
        self._n_streams = int(n_streams)
        self._device = config.CUDA_DEVICE if device is None else device
        self._filter.calculate_filter_blocks_fd(self._block_length)
        self._plan = None
        self._blocks_seen = 0  # blocks processed by the current plan (mono / full-width choice)
        if type(self) is OverlapSaveConvolver:
            # rows of `_input_block_td` = filter channels (`_convolver.py:477-480`); a mono input
            # is broadcast into all of them (`:588-591`), for which the plan has a cheaper form
            self._make_plan(self._filter.get_filter_blocks_fd().shape[-2])

    def _make_plan(self, n_in):
        hfd = self._filter.get_filter_blocks_fd()  # [P, 1, Cout, F]
        self._plan = OlsPlan(self._n_streams, self._block_length, n_in, hfd.shape[-2], hfd.shape[0],
                             device=self._device)
        self._plan.set_filter(hfd[:, 0])
        self._plan.set_passthrough(self._is_passthrough)
        self._blocks_seen = 0

    def __copy__(self):
        """Independent clone for ``JackRendererBenchmark.add_convolver`` (``_convolver.py:482-491``);
        unlike the reference's clones this one does not share signal buffers."""
        _filter = copy(self._filter)
        new = type(self)(_filter, self._block_length, n_streams=self._n_streams, device=self._device)
        new.set_passthrough(self._is_passthrough)
        return new

    def __str__(self):
        return (f"[{super().__str__()[1:-1]}, _block_length={self._block_length}, "
                f"_n_streams={self._n_streams}]")

    def _clear_buffers(self):
        if self._plan is not None:
            self._plan.reset()

    def _apply_passthrough(self):
        if self._plan is not None:
            self._plan.set_passthrough(self._is_passthrough)

    def filter_block(self, input_block_td):
        """One block through the partitioned convolver (``_convolver.py:524-571``).

        input [channels; B] or [1; B] (mono, broadcast) -> output [channels; B]; with
        ``n_streams > 1`` both carry a leading stream axis.  torch CUDA tensors are processed
        in place on the device and a device tensor is returned.
        """
        if input_block_td is None:
            return Convolver.filter_block(self, None)
        n_rows = input_block_td.shape[-2]
        if n_rows != self._plan.n_in:
            # The reference accepts a mono or a full-width block on every call and broadcasts mono
            # into the same buffers (`_convolver.py:588-591`).  Here the mono form has its own,
            # much cheaper plan (one input spectrum per block): it is selected by the FIRST block.
            # A mono block after full-width ones is broadcast on the way in (state kept, exactly
            # the reference's arithmetic); a full-width block on a mono plan cannot be honoured
            # without dropping the delay line and is refused.
            if n_rows == 1 and self._blocks_seen == 0:
                self._make_plan(1)
            elif n_rows == 1:
                if _is_torch_tensor(input_block_td):
                    input_block_td = input_block_td.expand(
                        *input_block_td.shape[:-2], self._plan.n_in, input_block_td.shape[-1]).contiguous()
                else:
                    input_block_td = np.ascontiguousarray(np.broadcast_to(
                        input_block_td, input_block_td.shape[:-2] + (self._plan.n_in, input_block_td.shape[-1])))
            else:
                raise ValueError(
                    f"expected {self._plan.n_in} input channel(s) (the convolver was started with "
                    f"{'a mono' if self._plan.n_in == 1 else 'a full-width'} block), got {n_rows}")
        self._blocks_seen += 1
        plan = self._plan
        if _is_torch_tensor(input_block_td):
            import torch

            x = input_block_td.reshape(plan.S, -1, plan.B)
            out = torch.empty((plan.S, plan.n_out, plan.B), dtype=torch.float32, device=x.device)
            plan.process_device(x.contiguous(), out, torch.cuda.current_stream(x.device))
            return out if input_block_td.dim() == 3 else out[0]
        x = np.asarray(input_block_td, dtype=np.float32)
        batched = x.ndim == 3
        out = plan.process_host(x.reshape(plan.S, -1, plan.B))
        return out if batched else out[0]


class AdjustableFdConvolver(OverlapSaveConvolver):
    """HRIR / BRIR switching renderer (``_convolver.py:672-975``): per source the filter pair of the
    direction given by the head yaw is applied, previous and current filters are cross-faded."""

    def __init__(self, filter_set, block_length, source_positions, shared_tracker_data,
                 n_streams=1, device=None):
        super().__init__(filter_set=filter_set, block_length=block_length, n_streams=n_streams,
                         device=device)
        self._is_crossfade = True
        if not source_positions:
            raise ValueError(
                "no virtual source positions (according to playback file channel count) given.")
        self._tracker_deg = shared_tracker_data
        self._sources_deg = np.array(source_positions, dtype=np.float16)
        # cos^2 windows, evaluated exactly like `_convolver.py:741-747`
        rdtype = self._filter.get_dirac_td().dtype
        w = np.arange(self._block_length, dtype=rdtype)
        self._window_out_td = np.square(np.cos(w * np.pi / (2 * (self._block_length - 1))))
        self._window_in_td = np.flip(self._window_out_td).copy()
        if type(self) is AdjustableFdConvolver:
            hfd = self._filter._irs_blocks_fd  # [P, directions, 2, F]
            self._plan = FdPlan(self._n_streams, block_length, self._sources_deg.shape[0],
                                hfd.shape[1], hfd.shape[0], device=self._device)
            self._plan.set_filter(hfd, self._window_in_td, self._window_out_td)
            self._plan.set_sources(self._sources_deg[:, 0].astype(np.float32),
                                   1 if self._filter._is_hrir else -1)

    def __copy__(self):
        new = type(self)(copy(self._filter), self._block_length,
                         [tuple(float(v) for v in row) for row in self._sources_deg],
                         self._tracker_deg, n_streams=self._n_streams, device=self._device)
        new.set_crossfade(self._is_crossfade)
        new.set_passthrough(self._is_passthrough)
        return new

    def filter_block(self, input_block_td, yaw_deg=None):
        """One block through the direction-switched renderer (``_convolver.py:780-836``):
        input [sources; B] -> [2; B]; batched and torch-tensor forms as for the SH renderer."""
        if input_block_td is None:
            return Convolver.filter_block(self, None)
        plan = self._plan
        if yaw_deg is None:
            azim = self._tracker_deg[HeadTracker.DataIndex.AZIM] if self._tracker_deg is not None else 0.0
            yaw_deg = np.full(plan.S, azim, dtype=np.float32)
        if _is_torch_tensor(input_block_td):
            import torch

            x = input_block_td.reshape(plan.S, plan.NS, plan.B).contiguous()
            if not _is_torch_tensor(yaw_deg):
                yaw_deg = torch.as_tensor(np.asarray(yaw_deg, dtype=np.float32), device=x.device)
            out = torch.empty((plan.S, 2, plan.B), dtype=torch.float32, device=x.device)
            plan.process_device(x, yaw_deg.contiguous(), out, torch.cuda.current_stream(x.device))
            return out if input_block_td.dim() == 3 else out[0]
        x = np.asarray(input_block_td, dtype=np.float32)
        out = plan.process_host(x.reshape(plan.S, plan.NS, plan.B), yaw_deg)
        return out if x.ndim == 3 else out[0]

    def set_crossfade(self, new_state=None):
        """``_convolver.py:868-892``: switching off also clears the previous-filter state."""
        self._is_crossfade = (not self._is_crossfade) if new_state is None else bool(new_state)
        if type(self) is AdjustableFdConvolver and self._plan is not None:
            self._plan.set_crossfade(self._is_crossfade)
        return self._is_crossfade

    def get_input_channel_count(self):
        return self._sources_deg.shape[0]

    def get_output_channel_count(self):
        return self._filter._irs_td.shape[1]


class AdjustableShConvolver(AdjustableFdConvolver):
    """Binaural renderer in the spherical-harmonics domain (``_convolver.py:978-1341``)."""

    def __init__(self, filter_set, block_length, source_positions, shared_tracker_data,
                 n_streams=1, device=None):
        super().__init__(filter_set, block_length, source_positions, shared_tracker_data,
                         n_streams=n_streams, device=device)
        if self._sources_deg.shape[0] > 1:
            raise NotImplementedError(
                "more then one virtual source position given, which is not supported yet.")
        self._sh_m = None
        self._sh_m_rev_id = None
        self._sh_cur_order = None
        self._sh_bases_weighted = None
        self._comp = None
        self._comp_arir_config = None
        self._comp_mrf_limit = None
        self._comp_nm_override = None
        self._hpcf_filter = None
        self._pre_renderer = None
        self._pre_renderer_irs = None

    def __copy__(self):
        """The reference raises here ("not tested", ``_convolver.py:1051-1060``); this returns an
        independent renderer with the same tables and switches."""
        new = type(self)(copy(self._filter), self._block_length,
                         [tuple(float(v) for v in self._sources_deg[0])], self._tracker_deg,
                         n_streams=self._n_streams, device=self._device)
        if self._plan is not None:
            new._install_tables(self._sh_m, self._sh_bases_weighted, self._comp,
                                self._comp_arir_config, self._comp_mrf_limit,
                                self._comp_nm_override, recompute=False)
            new.set_crossfade(self._is_crossfade)
            new.set_passthrough(self._is_passthrough)
            if self._sh_cur_order != self._filter._sh_max_order:
                new.update_sh_processing(self._sh_cur_order)
            if self._hpcf_filter is not None:
                new.set_hpcf(self._hpcf_filter)
            if self._pre_renderer_irs is not None:
                new.set_pre_renderer(self._pre_renderer_irs)
        return new

    def __str__(self):
        return (f"[{super().__str__()[1:-1]}, _sh_cur_order={self._sh_cur_order}, "
                f"_is_crossfade={self._is_crossfade}]")

    # -- set-up -------------------------------------------------------------------------------
    def prepare_sh_processing(self, input_sh_config, mrf_limit_db, compensation_type,
                              logger=None, comp_nm=None):
        """Everything that can be computed before real-time operation (``_convolver.py:1071-1120``).

        comp_nm (extension): ready-made compensation spectra [K; 1; F] multiplied into the HRIR
        coefficients instead of those ``Compensation.generate_by_type`` would design.
        """
        from ._compensation import Compensation

        comp = [compensation_type, Compensation.Type.MRF]
        self._install_tables(input_sh_config.sh_m, input_sh_config.sh_bases_weighted, comp,
                             input_sh_config.arir_config, mrf_limit_db, comp_nm, logger=logger)
        if config.IS_DEBUG_MODE:
            self._debug_filter_block(input_count=self._sh_bases_weighted.shape[-1])

    def _install_tables(self, sh_m, sh_bases_weighted, comp, arir_config, mrf_limit_db, comp_nm,
                        logger=None, recompute=True):
        order = self._filter._sh_max_order
        self._sh_m = np.array(sh_m).copy()
        self._sh_m_rev_id = sph.reverse_mn_ids(order)
        self._sh_cur_order = order
        self._sh_bases_weighted = np.array(sh_bases_weighted).copy()
        self._comp = comp
        self._comp_arir_config = arir_config
        self._comp_mrf_limit = mrf_limit_db
        self._comp_nm_override = comp_nm
        if recompute:
            self._apply_sh_compensation(is_plot=True, logger=logger, upload=False)
        hnm = self._filter.get_filter_blocks_nm()
        n_ch = self._sh_bases_weighted.shape[-1]
        self._plan = ShPlan(self._n_streams, self._block_length, n_ch, order, hnm.shape[-2],
                            hnm.shape[0], device=self._device)
        self._plan.set_tables(self._sh_bases_weighted, hnm, self._sh_m, self._sh_m_rev_id,
                              self._window_in_td, self._window_out_td)
        self._plan.set_source(float(self._sources_deg[0, 0]), 1 if self._filter._is_hrir else -1)
        self._plan.set_crossfade(self._is_crossfade)
        self._plan.set_passthrough(self._is_passthrough)
        if self._hpcf_filter is not None:
            self.set_hpcf(self._hpcf_filter)

    def update_sh_processing(self, sh_new_order, logger=None):
        """Change the rendering order at run time (``_convolver.py:1123-1168``)."""
        max_order = self._filter._sh_max_order
        sh_new_order = max_order if sh_new_order is None else int(sh_new_order)
        if not 0 <= sh_new_order <= max_order:
            msg = f'value "{sh_new_order}" out of bounds, "set SH processing order" ignored.'
            logger.error(msg) if logger else print(msg)
            return
        if self._sh_cur_order == sh_new_order:
            return sh_new_order
        was_running = config.IS_RUNNING.is_set()
        config.IS_RUNNING.clear()
        self._sh_cur_order = sh_new_order
        self._apply_sh_compensation(is_plot=False, logger=logger)
        self._plan.set_order(sh_new_order)
        if was_running:
            config.IS_RUNNING.set()
        return sh_new_order

    def _apply_sh_compensation(self, is_plot=True, logger=None, upload=True):
        """(Re)build `_irs_blocks_nm` = SH-transformed HRIR x compensation spectra
        (``_convolver.py:1171-1220``) and hand it to the device plan."""
        from ._compensation import Compensation

        Compensation.reset_config(is_plot=is_plot)
        if config.IS_SH_TRANSFORM_ON_DEVICE and type(self._filter) in (FilterSetMiro, FilterSetSofa):
            self._filter.calculate_filter_blocks_nm(device=self._device)
        else:
            self._filter.calculate_filter_blocks_nm()
        if self._comp_nm_override is not None:
            comp_nm = self._comp_nm_override
        else:
            comp_nm = Compensation.generate_by_type(
                compensation_types=self._comp, filter_set=self._filter,
                arir_config=self._comp_arir_config, amp_limit_db=self._comp_mrf_limit,
                nfft=None, nfft_padded=2 * self._block_length, logger=logger)
        self._filter._irs_blocks_nm *= comp_nm
        if upload and self._plan is not None:
            self._plan.update_filter(self._filter._irs_blocks_nm)

    # -- run time -----------------------------------------------------------------------------
    def _clear_buffers(self):
        if self._plan is not None:
            self._plan.reset()
        if self._pre_renderer is not None:
            self._pre_renderer._clear_buffers()

    def _apply_passthrough(self):
        if self._plan is not None:
            self._plan.set_passthrough(self._is_passthrough)

    def set_crossfade(self, new_state=None):
        """``_convolver.py:1318-1341``: switching off also forgets the previous rotation."""
        super().set_crossfade(new_state)
        if self._plan is not None:
            self._plan.set_crossfade(self._is_crossfade)
        return self._is_crossfade

    def get_input_channel_count(self):
        """Array channels once prepared (``_convolver.py:1222-1229``)."""
        if self._sh_bases_weighted is None:
            return super().get_input_channel_count()
        return self._sh_bases_weighted.shape[-1]

    def filter_block(self, input_block_td, yaw_deg=None):
        """Render one block (``_convolver.py:1231-1316``).

        input [channels; B] float32 -> [2; B] float32 (a fresh, writable array).  The head yaw is
        read from the shared tracker array (index ``AZIM``, degrees) unless ``yaw_deg`` is given.
        Batched: input [n_streams; channels; B] with ``yaw_deg`` [n_streams]; torch CUDA tensors
        stay on the device.
        """
        if input_block_td is None:
            return Convolver.filter_block(self, None)
        if self._plan is None:
            raise RuntimeError(self._filter._ERROR_MSG_NM)
        plan = self._plan
        if yaw_deg is None:
            azim = self._tracker_deg[HeadTracker.DataIndex.AZIM] if self._tracker_deg is not None else 0.0
            yaw_deg = np.full(plan.S, azim, dtype=np.float32)
        if self._pre_renderer is not None and input_block_td.shape[-2] == 1 and plan.C != 1:
            return self._filter_block_pre_rendered(input_block_td, yaw_deg)
        if _is_torch_tensor(input_block_td):
            import torch

            x = input_block_td.reshape(plan.S, plan.C, plan.B).contiguous()
            if not _is_torch_tensor(yaw_deg):
                yaw_deg = torch.as_tensor(np.asarray(yaw_deg, dtype=np.float32), device=x.device)
            out = torch.empty((plan.S, plan.E, plan.B), dtype=torch.float32, device=x.device)
            plan.process_device(x, yaw_deg.contiguous(), out, torch.cuda.current_stream(x.device))
            return out if input_block_td.dim() == 3 else out[0]
        x = np.asarray(input_block_td, dtype=np.float32)
        out = plan.process_host(x.reshape(plan.S, plan.C, plan.B), yaw_deg)
        return out if x.ndim == 3 else out[0]

    def set_pre_renderer(self, arir_irs_td):
        """Extension: fold the reference's ARIR pre-renderer into this renderer.

        The reference renders a mono source through an array room impulse response with a separate
        client (``_run.py:240-288``: ``OverlapSaveConvolver`` with a mono input and one filter per
        array channel, ``_convolver.py:588-591``) whose C output channels feed this renderer.  The
        array signals enter the SH chain only through the real analysis rows ``z = R x``
        (``R`` = real / imaginary parts of ``_sh_bases_weighted``, ``:1260-1263`` applied before the
        FFT), and the pre-renderer is linear and time invariant, so ``R`` folds into its filters
        once: ``z_j = (sum_c R[j, c] arir_c) * source``.  The pre-renderer then produces the J rows
        directly (169 instead of 434 channels for order 12 on a 434-channel array), and the
        renderer skips its analysis GEMM.  Same output up to fp32 reassociation.

        arir_irs_td: array impulse responses ``[channels; taps]`` (or a loaded array ``FilterSet``
        with ``_irs_td [1; channels; taps]``); ``None`` removes the pre-renderer.  Afterwards
        ``filter_block`` accepts the MONO source block ``[n_streams; 1; B]`` (or ``[1; B]``).
        """
        from ._filter_set import FilterSet

        self._pre_renderer = None
        self._pre_renderer_irs = None
        if arir_irs_td is None:
            return
        if self._plan is None:
            raise RuntimeError(self._filter._ERROR_MSG_NM)
        irs = getattr(arir_irs_td, "_irs_td", arir_irs_td)
        irs = np.asarray(irs, dtype=np.float64)
        irs = irs.reshape(-1, irs.shape[-1])
        if irs.shape[0] != self._plan.C:
            raise ValueError(f"the array impulse responses have {irs.shape[0]} channels, the renderer expects {self._plan.C}")
        rows = self._plan.analysis_rows().astype(np.float64)       # [J, C]
        combined = (rows @ irs).astype(np.float32)                 # [J, taps]: the ARIR in the analysis domain
        fs = FilterSet.create_instance_by_type(combined, "FIR_MULTICHANNEL")
        fs.load(block_length=self._block_length, is_single_precision=True, is_prevent_logging=True)
        pre = OverlapSaveConvolver(fs, self._block_length, n_streams=self._n_streams, device=self._device)
        pre._make_plan(1)  # mono source, J outputs (per-bin GEMM on tensor cores for >= 32 rows)
        self._pre_renderer = pre
        self._pre_renderer_irs = irs

    def _filter_block_pre_rendered(self, source_td, yaw_deg):
        """Mono source block(s) -> pre-renderer (writes the analysis rows into the plan) -> renderer."""
        import torch

        plan, pre = self._plan, self._pre_renderer._plan
        was_numpy = not _is_torch_tensor(source_td)
        dev = torch.device("cuda", self._device)
        x = torch.as_tensor(np.asarray(source_td, dtype=np.float32), device=dev) if was_numpy else source_td
        batched = x.dim() == 3
        x = x.reshape(plan.S, 1, plan.B).contiguous()
        if not _is_torch_tensor(yaw_deg):
            yaw_deg = torch.as_tensor(np.asarray(yaw_deg, dtype=np.float32).reshape(plan.S), device=dev)
        stream = torch.cuda.current_stream(dev)
        out = torch.empty((plan.S, plan.E, plan.B), dtype=torch.float32, device=dev)
        pre.process_device(x, plan.rows_buffer(), stream)
        plan.process_rows_device(yaw_deg.contiguous(), out, stream)
        if was_numpy:
            out = out.cpu().numpy()
        return out if batched else out[0]

    def set_hpcf(self, hpcf_filter_set):
        """Extension: run a headphone-compensation filter behind the renderer inside the renderer's
        own kernels.  The reference chains the HPCF as a separate client
        (``_run.py:290-338`` -> ``OverlapSaveConvolver.filter_block``, ``_convolver.py:524-571``);
        the result is the same, one launch and one HBM round trip of the binaural block less.
        `hpcf_filter_set`: a loaded 2-channel ``FilterSetMultiChannel`` (``None`` detaches)."""
        self._hpcf_filter = hpcf_filter_set
        if self._plan is None:
            return  # applied by `_install_tables` once the renderer is prepared
        if hpcf_filter_set is None:
            self._plan.set_hpcf(None)
            return
        hpcf_filter_set.calculate_filter_blocks_fd(self._block_length)
        hfd = hpcf_filter_set.get_filter_blocks_fd()  # [P, 1, 2, F] ([P, 2, 1, F] for an "HPCF_FIR" set)
        if hfd.shape[1] * hfd.shape[2] != 2:
            raise ValueError(f"the headphone compensation filter must have 2 channels, got {hfd.shape[1:3]}")
        self._plan.set_hpcf(hfd.reshape(hfd.shape[0], 2, hfd.shape[-1]))

    def submit_block(self, input_block_td, yaw_deg=None):
        """Extension: asynchronous ``filter_block`` for host blocks.  Returns a ticket whose
        ``result()`` yields what ``filter_block`` would have returned.  Blocks are rendered in
        submission order; up to three may be in flight, so a feeder that submits block t+1 before
        it collects block t overlaps the input copy with the rendering and the result copy (a
        real-time feeder does exactly that: the reference's JACK thread hands over one block per
        period, ``_jack_client.py:713-867``).  The input array must not be modified before
        ``result()`` returns."""
        if self._plan is None:
            raise RuntimeError(self._filter._ERROR_MSG_NM)
        plan = self._plan
        if yaw_deg is None:
            azim = self._tracker_deg[HeadTracker.DataIndex.AZIM] if self._tracker_deg is not None else 0.0
            yaw_deg = np.full(plan.S, azim, dtype=np.float32)
        x = np.asarray(input_block_td, dtype=np.float32)
        ticket = plan.submit_host(x.reshape(plan.S, plan.C, plan.B), yaw_deg)
        if x.ndim != 3:
            ticket._out = ticket._out[0]
        return ticket

