"""Block feeder: the receive / process / deliver cycle of a JACK client without JACK.

Mirror of the parts of ``JackClient`` (``ReTiSAR/_jack_client.py``) that surround the convolver in
the per-block callback (``:130-152``): input delay line (``_process_receive``, ``:713-747``),
``_process`` = ``convolver.filter_block`` (``JackRenderer._process``, ``_jack_renderer.py:296-315``),
output volume / mute, level metering and clipping detection (``_process_deliver``, ``:770-867``,
levels as ``tools.calculate_rms`` / ``calculate_peak``, ``tools.py:1117-1169``), and the dropout
counter (``xrun`` callback, ``:220-235``, read and reset like ``get_dropout_counter`` used by
``_run_benchmark.py:311-321``).  JACK itself (ports, server, connections) and OSC are out of scope:
blocks are handed in and out as arrays, and a dropout is a block whose processing took longer
than the block period (the condition under which JACK reports an xrun).

Blocks are ``[channels; block]`` numpy arrays like the reference's, or -- batched extension --
``[n_streams; channels; block]`` numpy arrays or torch CUDA tensors; with device tensors the output
stage runs in ``rtb_output_stage`` (one pass: scale in place, RMS, peak, clip count).
"""
import time

import numpy as np

from ._convolver import DelayBuffer, _is_torch_tensor


def calculate_rms(data, is_level=False):
    """``tools.calculate_rms`` (``tools.py:1117-1145``) for real time-domain data (last axis)."""
    rms = np.sqrt(np.mean(np.square(np.abs(data)), axis=-1))
    if is_level:
        rms = np.asarray(rms)
        rms[np.nonzero(rms == 0)] = np.nan  # prevent zeros
        with np.errstate(divide="ignore"):
            rms = 20 * np.log10(rms)
        rms[np.isnan(rms)] = -200  # zeros are reported as -200 dB
    return rms


def calculate_peak(data_td, is_level=False):
    """``tools.calculate_peak`` (``tools.py:1148-1169``)."""
    peak = np.nanmax(np.abs(data_td), axis=-1)
    if is_level:
        peak = np.asarray(peak)
        peak[np.nonzero(peak == 0)] = np.nan
        with np.errstate(divide="ignore"):
            peak = 20 * np.log10(peak)
        peak[np.isnan(peak)] = -200
    return peak


class BlockFeeder(object):
    """One client's callback around a convolver (or around nothing: straight passthrough like
    ``JackClient._process``, ``_jack_client.py:749-768``).

    Attributes mirror the reference's: ``_output_volume`` (linear), ``_output_volume_relative``
    ([channels; 1]), ``_output_mute``, ``_input_buffer`` (``DelayBuffer``), ``_is_measure_levels``,
    ``_is_detect_clipping``, ``_counter_dropout``.  ``last_rms`` / ``last_peak`` hold the levels
    of the last delivered block in dBFS (the reference sends them over OSC or logs them).
    """

    def __init__(self, convolver, block_length, sample_rate=48000, n_output_channels=None,
                 output_volume_dbfs=0.0, input_delay_ms=0.0, is_measure_levels=False,
                 is_detect_clipping=True, logger=None):
        self._convolver = convolver
        self._block_length = block_length
        self._sample_rate = sample_rate
        self._logger = logger
        self._is_measure_levels = is_measure_levels
        self._is_detect_clipping = is_detect_clipping
        self._output_mute = False
        self._output_volume = pow(10, output_volume_dbfs / 20.0)  # `_jack_client.py:102`
        n_out = n_output_channels
        if n_out is None and convolver is not None:
            n_out = convolver.get_output_channel_count()
        self._n_out = n_out
        self._output_volume_relative = None if n_out is None else np.ones((n_out, 1))
        self._input_buffer = DelayBuffer(sample_rate, block_length, delay_ms=input_delay_ms)
        self._counter_dropout = 0
        self._counter_clipping = 0
        self._block_period = block_length / sample_rate
        self.last_rms = None
        self.last_peak = None
        self.last_block_seconds = 0.0
        self._d_gain = self._d_levels = self._d_clip = None
        self._d_clip_seen = 0  # device clip count already added to `_counter_clipping`

    # ---- configuration (names of `_jack_client.py:560-711`) ----------------------------------
    def init_input_buffer(self, max_length_sec, channel_count, n_streams=None, device=None):
        """``self._input_buffer.init`` as done in ``JackClient.start`` (``:329-333``)."""
        self._input_buffer.init(max_length_sec, channel_count, n_streams=n_streams, device=device)

    def set_input_delay_ms(self, new_delay_ms=0.0):
        return self._input_buffer.set_delay(new_delay_ms)

    def set_output_mute(self, new_state=None):
        self._output_mute = (not self._output_mute) if new_state is None else bool(new_state)
        return self._output_mute

    def set_output_volume_db(self, value_db_fs=0.0):
        self._output_volume = 10 ** (float(value_db_fs) / 20.0)
        return value_db_fs

    def set_output_volume_port_relative_db(self, port, value_db_fs=0.0):
        port = int(port)
        if self._output_volume_relative is None or not 0 <= port < self._output_volume_relative.shape[0]:
            if self._logger:
                self._logger.error("provided port number is out of range, ignored.")
            return None
        self._output_volume_relative[port] = 10 ** (float(value_db_fs) / 20.0)
        return value_db_fs

    def get_dropout_counter(self, is_reset=True):
        """Dropouts since the last call (``JackClient.get_dropout_counter``)."""
        n = self._counter_dropout
        if is_reset:
            self._counter_dropout = 0
        return n

    # ---- the per-block callback --------------------------------------------------------------
    def _process_receive(self, input_td):
        return self._input_buffer.process_block(input_td)

    def _process(self, input_td, yaw_deg=None):
        if self._convolver is None:
            return input_td  # straight passthrough
        if yaw_deg is None:
            return self._convolver.filter_block(input_td)
        return self._convolver.filter_block(input_td, yaw_deg)

    def _process_deliver(self, output_td):
        if self._output_mute or output_td is None:
            return None  # the reference writes zeros to its ports
        if self._output_volume_relative is None:
            self._output_volume_relative = np.ones((output_td.shape[-2], 1))
        if _is_torch_tensor(output_td):
            return self._deliver_device(output_td)
        output_td *= (self._output_volume * self._output_volume_relative).astype(output_td.dtype)
        if self._is_measure_levels:
            self.last_rms = calculate_rms(output_td, is_level=True)
            self.last_peak = calculate_peak(output_td, is_level=True)
        if self._is_detect_clipping:
            n_clip = int((np.abs(output_td).max(axis=-1) > 1).sum())
            if n_clip and self._logger:
                self._logger.warning(f"output clipping detected ({n_clip} channel blocks).")
            self._counter_clipping += n_clip
        return output_td

    def _deliver_device(self, y):
        import torch

        from ._capi import check, lib
        from .plans import _dev_ptr, _stream_ptr

        if not y.is_contiguous():
            raise TypeError("expected a contiguous output tensor")
        n_ch, B = y.shape[-2], y.shape[-1]
        rows = y.numel() // B
        if self._d_gain is None or self._d_levels.shape[1] != rows:
            self._d_gain = torch.empty(n_ch, device=y.device, dtype=torch.float32)
            self._d_levels = torch.empty((2, rows), device=y.device, dtype=torch.float32)
            self._d_clip = torch.zeros(1, device=y.device, dtype=torch.int32)
            self._d_clip_seen = 0
            self._gain_host = None
        gain = (self._output_volume * self._output_volume_relative[:, 0]).astype(np.float32)
        if self._gain_host is None or not np.array_equal(gain, self._gain_host):
            self._d_gain.copy_(torch.from_numpy(gain))
            self._gain_host = gain
        measure = self._is_measure_levels
        check(lib.rtb_output_stage(_dev_ptr(y), _dev_ptr(self._d_gain),
                                   _dev_ptr(self._d_levels[0]) if measure else None,
                                   _dev_ptr(self._d_levels[1]) if measure else None,
                                   _dev_ptr_int(self._d_clip) if self._is_detect_clipping else None,
                                   rows // n_ch, n_ch, B, 1,
                                   _stream_ptr(torch.cuda.current_stream(y.device))))
        if measure:
            lv = self._d_levels.cpu().numpy().reshape((2,) + tuple(y.shape[:-1]))
            self.last_rms, self.last_peak = lv[0], lv[1]
            if self._is_detect_clipping:  # the count is read back together with the levels
                self._refresh_device_clipping()
        return y

    def _refresh_device_clipping(self):
        """Add what the device counted since the last read to the host counter (the host path adds
        to the same counter) and warn like ``_jack_client.py:853-862``."""
        if self._d_clip is None:
            return
        total = int(self._d_clip.item())
        delta = total - self._d_clip_seen
        self._d_clip_seen = total
        if delta > 0:
            self._counter_clipping += delta
            if self._logger:
                self._logger.warning(f"output clipping detected ({delta} channel blocks).")

    def get_clipping_counter(self):
        self._refresh_device_clipping()
        return self._counter_clipping

    def process_block(self, input_td, yaw_deg=None):
        """``process(frames)`` (``_jack_client.py:130-152``): receive, process, deliver; counts a
        dropout when the block took longer than its period."""
        t0 = time.perf_counter()
        out = self._process_deliver(self._process(self._process_receive(input_td), yaw_deg))
        if out is not None and _is_torch_tensor(out):
            import torch

            torch.cuda.current_stream(out.device).synchronize()
        self.last_block_seconds = time.perf_counter() - t0
        if self.last_block_seconds > self._block_period and not self._output_mute:
            self._counter_dropout += 1
        return out


def _dev_ptr_int(t):
    import ctypes

    return ctypes.c_void_p(t.data_ptr())
