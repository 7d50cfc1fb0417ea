"""ctypes binding of ``libretisar_b200.so`` (C ABI declared in ``include/retisar_b200.h``).

There is deliberately no fallback: if the CUDA library has not been built, importing this module
raises.  Build it with ``python -c "import __graft_entry__ as g; g.build()"`` or
``retisar_b200/csrc/build.sh``.
"""
import ctypes
import os

# RTB_LIB_PATH: load another build of the same library (A/B runs of kernel variants)
_LIB_PATH = os.environ.get("RTB_LIB_PATH") or os.path.join(
    os.path.dirname(os.path.abspath(__file__)), "lib", "libretisar_b200.so")

RTB_OK, RTB_ERR_INVALID, RTB_ERR_STATE, RTB_ERR_UNSUPPORTED, RTB_ERR_CUDA, RTB_ERR_NOMEM = range(6)
RTB_Q_SYMMETRIC, RTB_Q_ROWS, RTB_Q_SIGNALS, RTB_Q_LAUNCHES, RTB_Q_STATE_BYTES = 1, 2, 3, 4, 5
RTB_Q_ANALYSIS_NS, RTB_Q_RENDER_NS, RTB_Q_TENSOR_CORE, RTB_Q_RENDER_KERNEL = 6, 7, 8, 9
RTB_Q_GRAPH_REPLAYS = 10
RENDER_KERNEL_NAMES = {0: "sh_render_kernel", 1: "sh_render_mg_kernel", 2: "sh_render_split_kernel",
                       3: "sh_partition_mac2_kernel", 4: "sh_passthrough_kernel", 6: "sh_partition_tc_kernel"}

# every symbol include/retisar_b200.h declares: name -> (restype, argtypes)
_c_float_p = ctypes.POINTER(ctypes.c_float)
_vp = ctypes.c_void_p
_i = ctypes.c_int
SIGNATURES = {
    "rtb_version": (_i, []),
    "rtb_last_error": (ctypes.c_char_p, []),
    "rtb_device_count": (_i, []),
    "rtb_sh_plan_create": (_i, [ctypes.POINTER(_vp), _i, _i, _i, _i, _i, _i, _i]),
    "rtb_sh_plan_set_tables": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rtb_sh_plan_update_filter": (_i, [_vp, _vp]),
    "rtb_sh_plan_set_source": (_i, [_vp, ctypes.c_float, _i]),
    "rtb_sh_plan_set_order": (_i, [_vp, _i]),
    "rtb_sh_plan_set_passthrough": (_i, [_vp, _i]),
    "rtb_sh_plan_set_crossfade": (_i, [_vp, _i]),
    "rtb_sh_plan_reset": (_i, [_vp]),
    "rtb_sh_plan_set_hpcf": (_i, [_vp, _vp, _i]),
    "rtb_sh_plan_get_analysis_rows": (_i, [_vp, _vp]),
    "rtb_sh_plan_rows_buffer": (_i, [_vp, ctypes.POINTER(_vp)]),
    "rtb_sh_process_block_rows": (_i, [_vp, _vp, _vp, _vp]),
    "rtb_sh_transform_filter": (_i, [_i, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "rtb_sh_process_block": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "rtb_sh_process_blocks": (_i, [_vp, _vp, _vp, _vp, _i, ctypes.c_longlong, _i, _vp]),
    "rtb_sh_process_block_host": (_i, [_vp, _vp, _vp, _vp]),
    "rtb_sh_submit_block_host": (_i, [_vp, _vp, _vp, _vp, ctypes.POINTER(ctypes.c_longlong)]),
    "rtb_sh_wait_block_host": (_i, [_vp, ctypes.c_longlong]),
    "rtb_sh_plan_query": (_i, [_vp, _i, ctypes.POINTER(ctypes.c_int64)]),
    "rtb_sh_plan_set_profiling": (_i, [_vp, _i]),
    "rtb_sh_plan_destroy": (_i, [_vp]),
    "rtb_ols_plan_create": (_i, [ctypes.POINTER(_vp), _i, _i, _i, _i, _i, _i]),
    "rtb_ols_plan_set_filter": (_i, [_vp, _vp]),
    "rtb_ols_plan_set_passthrough": (_i, [_vp, _i]),
    "rtb_ols_plan_reset": (_i, [_vp]),
    "rtb_ols_process_block": (_i, [_vp, _vp, _vp, _vp]),
    "rtb_ols_process_block_host": (_i, [_vp, _vp, _vp]),
    "rtb_ols_plan_query": (_i, [_vp, _i, ctypes.POINTER(ctypes.c_int64)]),
    "rtb_ols_plan_set_profiling": (_i, [_vp, _i]),
    "rtb_ols_plan_destroy": (_i, [_vp]),
    "rtb_fd_plan_create": (_i, [ctypes.POINTER(_vp), _i, _i, _i, _i, _i, _i]),
    "rtb_fd_plan_set_filter": (_i, [_vp, _vp, _vp, _vp]),
    "rtb_fd_plan_set_sources": (_i, [_vp, _vp, _i]),
    "rtb_fd_plan_set_passthrough": (_i, [_vp, _i]),
    "rtb_fd_plan_set_crossfade": (_i, [_vp, _i]),
    "rtb_fd_plan_reset": (_i, [_vp]),
    "rtb_fd_process_block": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "rtb_fd_process_block_host": (_i, [_vp, _vp, _vp, _vp]),
    "rtb_fd_plan_query": (_i, [_vp, _i, ctypes.POINTER(ctypes.c_int64)]),
    "rtb_fd_plan_destroy": (_i, [_vp]),
    "rtb_host_alloc": (_i, [ctypes.POINTER(_vp), ctypes.c_ulonglong, _i]),
    "rtb_host_free": (_i, [_vp]),
    "rtb_output_stage": (_i, [_vp, _vp, _vp, _vp, _vp, ctypes.c_longlong, _i, _i, _i, _vp]),
}


def library_path():
    return _LIB_PATH


def _load():
    if not os.path.isfile(_LIB_PATH):
        raise RuntimeError(
            f"CUDA library {_LIB_PATH} has not been built (no CPU fallback exists); run "
            f"`python -c \"import __graft_entry__ as g; g.build()\"` or retisar_b200/csrc/build.sh"
        )
    lib = ctypes.CDLL(_LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


lib = _load()

_EXC = {
    RTB_ERR_INVALID: ValueError,
    RTB_ERR_STATE: RuntimeError,
    RTB_ERR_UNSUPPORTED: NotImplementedError,
    RTB_ERR_CUDA: RuntimeError,
    RTB_ERR_NOMEM: MemoryError,
}


def check(rc):
    """Map an RTB status code to the exception type the reference raises for that condition."""
    if rc != RTB_OK:
        msg = lib.rtb_last_error().decode("utf-8", "replace")
        raise _EXC.get(rc, RuntimeError)(msg)


def np_ptr(a):
    """Pointer to the data of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return ctypes.c_void_p(a.ctypes.data)
