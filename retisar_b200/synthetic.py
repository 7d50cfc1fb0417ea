"""Seeded synthetic inputs shaped like the reference's data.

Used by the parity tests, ``bench.py`` and ``oracle/make_golden.py``.  Plays the role of the
reference's ``tools.generate_noise`` (``ReTiSAR/tools.py:999-1033``) and of its (absent) measured
HRIR / array data sets: same shapes and scales, deterministic content.
"""
import os

import numpy as np

_GRID_FILE = os.path.join(os.path.dirname(__file__), "data", "array_grids.npz")


def load_array_grid(name):
    """Return (azimuth[rad], colatitude[rad], weight) of a measured array geometry.

    ``em32`` = Eigenmike em32 (32 ch), ``le110`` = Lebedev 110; values extracted once from the
    MIRO structs shipped with the reference (``res/ARIR/RT_calib_EM32ch_struct.mat``,
    ``res/ARIR/DRIR_sim_LE110_PW_struct.mat``) by ``oracle/make_golden.py``.
    """
    with np.load(_GRID_FILE) as z:
        return z[f"{name}_azimuth"], z[f"{name}_colatitude"], z[f"{name}_weight"]


def random_sphere_grid(n_points, seed):
    """Pseudo-random (uniform on the sphere) sampling grid without quadrature weights."""
    rng = np.random.default_rng(seed)
    az = rng.uniform(0.0, 2.0 * np.pi, n_points)
    co = np.arccos(rng.uniform(-1.0, 1.0, n_points))
    return az, co


def decaying_noise_irs(shape, seed, scale=0.05, tau=None, dtype=np.float32):
    """Impulse responses ``rng.standard_normal * scale * exp(-n / tau)`` along the last axis."""
    rng = np.random.default_rng(seed)
    taps = shape[-1]
    tau = tau if tau else max(taps / 6.0, 1.0)
    env = np.exp(-np.arange(taps) / tau)
    return (rng.standard_normal(shape) * scale * env).astype(dtype)


def noise_blocks(n_blocks, n_ch, block_length, seed, scale=0.1, dtype=np.float32):
    """Input signal blocks [n_blocks; n_ch; block_length] ~ N(0, scale^2)."""
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((n_blocks, n_ch, block_length)) * scale).astype(dtype)


def yaw_walk(n_blocks, seed, start=None, step=2.0):
    """Head yaw per block in degrees (float32), a bounded random walk like a slow head turn."""
    rng = np.random.default_rng(seed)
    y0 = rng.uniform(0.0, 360.0) if start is None else start
    inc = rng.uniform(-step, step, n_blocks)
    inc[0] = 0.0
    return (y0 + np.cumsum(inc)).astype(np.float32)


YAW_EDGE_CASES = np.array(
    [0.0, 10.3, 123.456, 359.9, -0.1, 720.3, 45.0, 45.0, 200.0, -170.25, 359.99, 180.0,
     -359.9, 90.0, 270.5, 0.124],
    dtype=np.float32,
)
