"""Spherical-harmonics helpers the render path needs.

Replacements for the few functions of the third-party package ``sound_field_analysis`` that the
reference calls on this path (``sph.mnArrays``, ``sph.reverseMnIds``, ``sph.sph_harm_all`` -- call
sites ``ReTiSAR/_filter_set.py:1259,1263`` and ``_convolver.py:1095``), built on scipy.
"""
import numpy as np
from scipy.special import sph_harm_y


def mn_arrays(order):
    """Stacked SH orders m and degrees n: m = [0, -1, 0, 1, -2, ...], n = [0, 1, 1, 1, 2, ...]."""
    n = np.repeat(np.arange(order + 1), 2 * np.arange(order + 1) + 1)
    m = np.arange(n.size) - n * n - n
    return m, n


def reverse_mn_ids(order):
    """For every stacked coefficient (n, m) the index of (n, -m)."""
    m, n = mn_arrays(order)
    return (n * n + n - m).tolist()


def sph_harm_all(order, azimuth, colatitude):
    """Complex spherical harmonics on a sampling grid, size [points; (order + 1)^2]."""
    m, n = mn_arrays(order)
    az = np.asarray(azimuth, dtype=np.float64).reshape(-1, 1)
    co = np.asarray(colatitude, dtype=np.float64).reshape(-1, 1)
    return sph_harm_y(n[np.newaxis, :], m[np.newaxis, :], co, az)
