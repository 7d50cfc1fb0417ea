"""Data contract of the head tracker (reference ``ReTiSAR/_tracker.py:73-81``).

The render path reads index ``AZIM`` of a 6-float shared array, in degrees
(``_convolver.py:961``).  Serial-port tracker drivers are out of scope.
"""
import multiprocessing
from enum import IntEnum


class HeadTracker:
    class DataIndex(IntEnum):
        X, Y, Z, AZIM, ELEV, TILT = range(6)

    @staticmethod
    def create_shared_array():
        """`multiprocessing.Array('f', 6)` as the reference tracker allocates it (`:156-164`)."""
        return multiprocessing.get_context("fork").Array("f", 6)
