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

    def analysis_rows(self):
        """The real analysis matrix R [J, C] of the plan (z = R x; pre-rendered mode)."""
        rows = np.empty((self.query(_capi.RTB_Q_ROWS), self.C), dtype=np.float32)
        check(lib.rtb_sh_plan_get_analysis_rows(self._h, np_ptr(rows)))
        return rows

    def rows_buffer(self):
        """Device address of the [S, J, B] buffer that must receive the next block's analysis rows."""
        ptr = ctypes.c_void_p()
        check(lib.rtb_sh_plan_rows_buffer(self._h, ctypes.byref(ptr)))
        return ptr.value

    def process_rows_device(self, d_yaw, d_out, stream=None):
        """Render the block whose analysis rows were written into `rows_buffer()`."""
        check(lib.rtb_sh_process_block_rows(self._h, _dev_ptr(d_yaw), _dev_ptr(d_out), _stream_ptr(stream)))

    def set_hpcf(self, hfd):
        """Attach a 2-channel headphone-compensation filter behind the renderer: hfd [P, 2, F]
        complex (``FilterSet.get_filter_blocks_fd()[:, 0]``); ``None`` detaches it."""
        if hfd is None:
            check(lib.rtb_sh_plan_set_hpcf(self._h, None, 0))
            return
        hfd = np.ascontiguousarray(hfd, dtype=np.complex64)
        if hfd.ndim != 3 or hfd.shape[1:] != (2, self.F):
            raise ValueError(f"HPCF filter blocks shape {hfd.shape} != (P, 2, {self.F})")
        check(lib.rtb_sh_plan_set_hpcf(self._h, np_ptr(hfd), hfd.shape[0]))

    def query(self, what):
        v = ctypes.c_int64()
        check(lib.rtb_sh_plan_query(self._h, what, ctypes.byref(v)))
        return v.value

    def set_profiling(self, on):
        check(lib.rtb_sh_plan_set_profiling(self._h, int(bool(on))))

    def _host_args(self, x, yaw_deg, out):
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.shape != (self.S, self.C, self.B):
            raise ValueError(f"input shape {x.shape} != {(self.S, self.C, self.B)}")
        yaw = None if yaw_deg is None else np.ascontiguousarray(yaw_deg, dtype=np.float32).reshape(self.S)
        if out is None:
            out = _OUT_POOL.take((self.S, self.E, self.B))
        return x, yaw, out

    def process_host(self, x, yaw_deg, out=None):
        """x [S,C,B] float32 host, yaw_deg [S] float32 host -> out [S,E,B] float32 host (a fresh,
        writable, page-locked array the caller may keep or modify in place)."""
        x, yaw, out = self._host_args(x, yaw_deg, out)
        check(lib.rtb_sh_process_block_host(self._h, np_ptr(x), np_ptr(yaw), np_ptr(out)))
        return out

    def submit_host(self, x, yaw_deg, out=None):
        """Asynchronous form of `process_host`: returns a ticket object whose ``result()`` blocks
        until the ear signals have arrived.  Up to three blocks may be in flight; the input copy
        of a block overlaps the kernels and the result copy of the one before it."""
        x, yaw, out = self._host_args(x, yaw_deg, out)
        ticket = ctypes.c_longlong()
        check(lib.rtb_sh_submit_block_host(self._h, np_ptr(x), np_ptr(yaw), np_ptr(out),
                                           ctypes.byref(ticket)))
        return HostTicket(self, ticket.value, out, (x, yaw))

    def process_device(self, d_x, d_yaw, d_out, stream=None):
        """Asynchronous: device tensors / pointers [S,C,B], [S] -> [S,E,B]."""
        check(lib.rtb_sh_process_block(self._h, _dev_ptr(d_x), _dev_ptr(d_yaw), _dev_ptr(d_out),
                                       _stream_ptr(stream)))


    def process_blocks_device(self, d_x_ring, d_yaw_ring, d_out, first, count, stream=None):
        """`count` consecutive blocks enqueued by one library call: block i reads ring slot
        i % len(ring) of d_x_ring [R,S,C,B] / d_yaw_ring [R,S]; every block writes d_out."""
        n_ring = int(d_x_ring.shape[0])
        check(lib.rtb_sh_process_blocks(self._h, _dev_ptr(d_x_ring), _dev_ptr(d_yaw_ring),
                                        _dev_ptr(d_out), n_ring, int(first), int(count),
                                        _stream_ptr(stream)))


class HostTicket:
    """A block submitted through `ShPlan.submit_host`; keeps the input arrays alive until the
    result has been fetched (the library reads them asynchronously)."""

    def __init__(self, plan, ticket, out, keep):
        self._plan, self._ticket, self._out, self._keep = plan, ticket, out, keep

    def result(self):
        if self._keep is not None:
            check(lib.rtb_sh_wait_block_host(self._plan._h, self._ticket))
            self._keep = None
        return self._out


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
            out = _OUT_POOL.take((self.S, self.n_out, self.B))
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
            out = _OUT_POOL.take((self.S, 2, self.B))
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
            try:
                lib.rtb_host_free(self._ptr)
            except Exception:  # interpreter shutdown: the library may be gone already
                pass
            self._ptr = ctypes.c_void_p()

    __del__ = close


class PinnedPool:
    """Page-locked result buffers for the host entry points.

    `filter_block` must return a fresh, writable array that does not alias internal state (the
    reference's callers scale it in place, ``_jack_client.py:795``).  A new pageable ``np.empty`` per
    block costs an mmap / page-fault / munmap cycle and a staged (not direct) device-to-host copy.
    The pool hands out views of page-locked buffers instead and takes a buffer back once the
    caller has dropped every reference to it (``sys.getrefcount`` of the buffer's root array: views
    handed out -- and views of those views -- keep the root alive and counted).  A caller that
    keeps many results simply makes the pool grow, up to `max_bytes`; beyond that, and for tiny
    blocks, plain ``np.empty`` arrays are returned.
    """

    def __init__(self, max_bytes=1 << 30, min_bytes=1 << 16):
        self._roots = {}  # nbytes -> [root arrays (1-D uint8 views of page-locked memory)]
        self._bufs = []   # PinnedBuffer owners
        self._bytes = 0
        self.max_bytes, self.min_bytes = max_bytes, min_bytes

    def take(self, shape, dtype=np.float32):
        import sys

        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        if n < self.min_bytes:
            return np.empty(shape, dtype=dtype)
        roots = self._roots.setdefault(n, [])
        for root in roots:
            if sys.getrefcount(root) == 3:  # the list, the loop variable, getrefcount's argument
                return root.view(dtype).reshape(shape)
        if self._bytes + n > self.max_bytes:
            return np.empty(shape, dtype=dtype)
        try:
            buf = PinnedBuffer((n,), np.uint8)
        except (MemoryError, RuntimeError):
            return np.empty(shape, dtype=dtype)
        self._bufs.append(buf)
        self._bytes += n
        root = buf.array
        while isinstance(root.base, np.ndarray):  # numpy parents every view to the array over the raw buffer
            root = root.base
        buf.array = None  # the pool's list holds the only long-lived reference to the root
        roots.append(root)
        return root.view(dtype).reshape(shape)


_OUT_POOL = PinnedPool()
