"""Thin object wrappers over the C ABI plans (``include/retisar_b200.h``).

``ShPlan`` / ``OlsPlan`` own one ``rtb_*_plan`` handle each.  They accept host numpy arrays (the
reference's ``filter_block`` shape) or device pointers / torch CUDA tensors (batched streams).
The reference-shaped classes in ``_convolver.py`` are built on top of these.
"""
import ctypes

import numpy as np

from . import _capi
from ._capi import check, lib, np_ptr


def _dev_ptr(t):
    """Device pointer of a torch CUDA tensor (contiguous float32) or a raw integer address."""
    if isinstance(t, int):
        return ctypes.c_void_p(t)
    if t is None:
        return None
    if not t.is_cuda or not t.is_contiguous() or str(t.dtype) != "torch.float32":
        raise TypeError("expected a contiguous float32 CUDA tensor")
    return ctypes.c_void_p(t.data_ptr())


def _stream_ptr(stream):
    if stream is None:
        return None
    if isinstance(stream, int):
        return ctypes.c_void_p(stream)
    return ctypes.c_void_p(stream.cuda_stream)  # torch.cuda.Stream


class ShPlan:
    """S batched SH-domain render streams (C ABI ``rtb_sh_*``)."""

    def __init__(self, n_streams, block_length, n_channels, sh_max_order, n_ears=2,
                 n_partitions=1, device=0):
        self._h = ctypes.c_void_p()
        check(lib.rtb_sh_plan_create(ctypes.byref(self._h), device, n_streams, block_length,
                                     n_channels, sh_max_order, n_ears, n_partitions))
        self.S, self.B, self.C, self.E = n_streams, block_length, n_channels, n_ears
        self.order, self.P, self.device = sh_max_order, n_partitions, device
        self.K = (sh_max_order + 1) ** 2
        self.F = block_length + 1

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib.rtb_sh_plan_destroy(self._h)
            self._h = ctypes.c_void_p()

    __del__ = close

    def _hnm(self, hnm):
        hnm = np.ascontiguousarray(hnm, dtype=np.complex64)
        if hnm.shape != (self.P, self.K, self.E, self.F):
            raise ValueError(f"hnm shape {hnm.shape} != {(self.P, self.K, self.E, self.F)}")
        return hnm

    def set_tables(self, yw, hnm, sh_m, rev, win_in=None, win_out=None):
        yw = np.ascontiguousarray(yw, dtype=np.complex64)
        if yw.shape != (self.K, self.C):
            raise ValueError(f"sh_bases_weighted shape {yw.shape} != {(self.K, self.C)}")
        hnm = self._hnm(hnm)
        sh_m = np.ascontiguousarray(sh_m, dtype=np.int16)
        rev = np.ascontiguousarray(rev, dtype=np.int32)
        if sh_m.shape != (self.K,) or rev.shape != (self.K,):
            raise ValueError("sh_m / reverse ids must have (order + 1)^2 entries")
        if win_in is not None:
            win_in = np.ascontiguousarray(win_in, dtype=np.float32)
            win_out = np.ascontiguousarray(win_out, dtype=np.float32)
        check(lib.rtb_sh_plan_set_tables(self._h, np_ptr(yw), np_ptr(hnm), np_ptr(sh_m),
                                         np_ptr(rev), np_ptr(win_in), np_ptr(win_out)))

    def update_filter(self, hnm):
        check(lib.rtb_sh_plan_update_filter(self._h, np_ptr(self._hnm(hnm))))

    def set_source(self, source_azim_deg=0.0, tracker_dir=1):
        check(lib.rtb_sh_plan_set_source(self._h, float(source_azim_deg), int(tracker_dir)))

    def set_order(self, order):
        check(lib.rtb_sh_plan_set_order(self._h, int(order)))

    def set_passthrough(self, on):
        check(lib.rtb_sh_plan_set_passthrough(self._h, int(bool(on))))

    def set_crossfade(self, on):
        check(lib.rtb_sh_plan_set_crossfade(self._h, int(bool(on))))

    def reset(self):
        check(lib.rtb_sh_plan_reset(self._h))

    def query(self, what):
        v = ctypes.c_int64()
        check(lib.rtb_sh_plan_query(self._h, what, ctypes.byref(v)))
        return v.value

    def set_profiling(self, on):
        check(lib.rtb_sh_plan_set_profiling(self._h, int(bool(on))))

    def process_host(self, x, yaw_deg, out=None):
        """x [S,C,B] float32 host, yaw_deg [S] float32 host -> out [S,E,B] float32 host."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.shape != (self.S, self.C, self.B):
            raise ValueError(f"input shape {x.shape} != {(self.S, self.C, self.B)}")
        yaw = None if yaw_deg is None else np.ascontiguousarray(yaw_deg, dtype=np.float32).reshape(self.S)
        if out is None:
            out = np.empty((self.S, self.E, self.B), dtype=np.float32)
        check(lib.rtb_sh_process_block_host(self._h, np_ptr(x), np_ptr(yaw), np_ptr(out)))
        return out

    def process_device(self, d_x, d_yaw, d_out, stream=None):
        """Asynchronous: device tensors / pointers [S,C,B], [S] -> [S,E,B]."""
        check(lib.rtb_sh_process_block(self._h, _dev_ptr(d_x), _dev_ptr(d_yaw), _dev_ptr(d_out),
                                       _stream_ptr(stream)))


class OlsPlan:
    """S batched partitioned overlap-save FIR convolvers (C ABI ``rtb_ols_*``)."""

    def __init__(self, n_streams, block_length, n_in, n_out, n_partitions, device=0):
        self._h = ctypes.c_void_p()
        check(lib.rtb_ols_plan_create(ctypes.byref(self._h), device, n_streams, block_length,
                                      n_in, n_out, n_partitions))
        self.S, self.B, self.n_in, self.n_out, self.P = n_streams, block_length, n_in, n_out, n_partitions
        self.F = block_length + 1
        self.device = device

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib.rtb_ols_plan_destroy(self._h)
            self._h = ctypes.c_void_p()

    __del__ = close

    def set_filter(self, hfd):
        hfd = np.ascontiguousarray(hfd, dtype=np.complex64)
        if hfd.shape != (self.P, self.n_out, self.F):
            raise ValueError(f"filter blocks shape {hfd.shape} != {(self.P, self.n_out, self.F)}")
        check(lib.rtb_ols_plan_set_filter(self._h, np_ptr(hfd)))

    def set_passthrough(self, on):
        check(lib.rtb_ols_plan_set_passthrough(self._h, int(bool(on))))

    def reset(self):
        check(lib.rtb_ols_plan_reset(self._h))

    def query(self, what):
        v = ctypes.c_int64()
        check(lib.rtb_ols_plan_query(self._h, what, ctypes.byref(v)))
        return v.value

    def set_profiling(self, on):
        check(lib.rtb_ols_plan_set_profiling(self._h, int(bool(on))))

    def process_host(self, x, out=None):
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.shape != (self.S, self.n_in, self.B):
            raise ValueError(f"input shape {x.shape} != {(self.S, self.n_in, self.B)}")
        if out is None:
            out = np.empty((self.S, self.n_out, self.B), dtype=np.float32)
        check(lib.rtb_ols_process_block_host(self._h, np_ptr(x), np_ptr(out)))
        return out

    def process_device(self, d_x, d_out, stream=None):
        check(lib.rtb_ols_process_block(self._h, _dev_ptr(d_x), _dev_ptr(d_out), _stream_ptr(stream)))


class FdPlan:
    """S batched HRIR / BRIR switching renderers (C ABI ``rtb_fd_*``)."""

    def __init__(self, n_streams, block_length, n_sources, n_dirs, n_partitions, device=0):
        self._h = ctypes.c_void_p()
        check(lib.rtb_fd_plan_create(ctypes.byref(self._h), device, n_streams, block_length,
                                     n_sources, n_dirs, n_partitions))
        self.S, self.B, self.NS, self.n_dirs, self.P = n_streams, block_length, n_sources, n_dirs, n_partitions
        self.F, self.E, self.device = block_length + 1, 2, device

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib.rtb_fd_plan_destroy(self._h)
            self._h = ctypes.c_void_p()

    __del__ = close

    def set_filter(self, hfd, win_in=None, win_out=None):
        hfd = np.ascontiguousarray(hfd, dtype=np.complex64)
        if hfd.shape != (self.P, self.n_dirs, 2, self.F):
            raise ValueError(f"filter blocks shape {hfd.shape} != {(self.P, self.n_dirs, 2, self.F)}")
        if win_in is not None:
            win_in = np.ascontiguousarray(win_in, dtype=np.float32)
            win_out = np.ascontiguousarray(win_out, dtype=np.float32)
        check(lib.rtb_fd_plan_set_filter(self._h, np_ptr(hfd), np_ptr(win_in), np_ptr(win_out)))

    def set_sources(self, source_azim_deg, tracker_dir=1):
        src = np.ascontiguousarray(source_azim_deg, dtype=np.float32).reshape(self.NS)
        check(lib.rtb_fd_plan_set_sources(self._h, np_ptr(src), int(tracker_dir)))

    def set_passthrough(self, on):
        check(lib.rtb_fd_plan_set_passthrough(self._h, int(bool(on))))

    def set_crossfade(self, on):
        check(lib.rtb_fd_plan_set_crossfade(self._h, int(bool(on))))

    def reset(self):
        check(lib.rtb_fd_plan_reset(self._h))

    def query(self, what):
        v = ctypes.c_int64()
        check(lib.rtb_fd_plan_query(self._h, what, ctypes.byref(v)))
        return v.value

    def process_host(self, x, yaw_deg, out=None):
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.shape != (self.S, self.NS, self.B):
            raise ValueError(f"input shape {x.shape} != {(self.S, self.NS, self.B)}")
        yaw = None if yaw_deg is None else np.ascontiguousarray(yaw_deg, dtype=np.float32).reshape(self.S)
        if out is None:
            out = np.empty((self.S, 2, self.B), dtype=np.float32)
        check(lib.rtb_fd_process_block_host(self._h, np_ptr(x), np_ptr(yaw), np_ptr(out)))
        return out

    def process_device(self, d_x, d_yaw, d_out, stream=None):
        check(lib.rtb_fd_process_block(self._h, _dev_ptr(d_x), _dev_ptr(d_yaw), _dev_ptr(d_out),
                                       _stream_ptr(stream)))


class PinnedBuffer:
    """Page-locked host block buffer from ``rtb_host_alloc``; ``.array`` is a numpy view.  Blocks
    handed to ``filter_block`` / ``process_host`` from such a buffer are copied to the device by a
    single DMA.  ``write_combined=True`` is for input blocks only (slow to read on the host)."""

    def __init__(self, shape, dtype=np.float32, write_combined=False):
        self.shape, self.dtype = tuple(shape), np.dtype(dtype)
        n = int(np.prod(self.shape)) * self.dtype.itemsize
        self._ptr = ctypes.c_void_p()
        check(lib.rtb_host_alloc(ctypes.byref(self._ptr), n, int(bool(write_combined))))
        buf = (ctypes.c_char * n).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=self.dtype).reshape(self.shape)

    def close(self):
        if getattr(self, "_ptr", None) is not None and self._ptr.value:
            self.array = None
            lib.rtb_host_free(self._ptr)
            self._ptr = ctypes.c_void_p()

    __del__ = close
