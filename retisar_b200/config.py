"""Run-time switches the render path reads (subset of the reference's ``ReTiSAR/config.py``).

Only the knobs the hot path consults exist here: ``IS_RUNNING`` (``config.py:321``),
``IS_SINGLE_PRECISION`` (``:269``), ``IS_DEBUG_MODE`` (``:316``).  ``IS_PYFFTW_MODE`` (``:291``) is
accepted for compatibility and ignored: the transforms run in the CUDA kernels.
"""
import multiprocessing

IS_SINGLE_PRECISION = True
IS_DEBUG_MODE = False
IS_PYFFTW_MODE = False
IS_RUNNING = multiprocessing.get_context("fork").Event()
IS_RUNNING.set()

CUDA_DEVICE = 0
"""CUDA device index new convolvers are placed on (no reference counterpart)."""

IS_SH_TRANSFORM_ON_DEVICE = False
"""Evaluate the SH transform of the HRIR set (``FilterSetMiro.calculate_filter_blocks_nm``) on the
tensor-core GEMM instead of numpy when the renderer (re)builds its tables (no reference
counterpart).  Off by default: the numpy product reproduces the reference's tables bit for bit,
the device product agrees to about 2e-6 relative."""
