"""B200-native per-block binaural rendering chain with the ReTiSAR convolver API.

Drop-in for the classes ``ReTiSAR._convolver`` / ``ReTiSAR._filter_set`` export; the arithmetic
runs in hand-written sm_100a CUDA kernels behind a C ABI (``include/retisar_b200.h``).
"""
__version__ = "0.1.0"

__all__ = ["config", "Convolver", "FilterSet", "HeadTracker", "Compensation"]

from . import config  # noqa: E402
from ._filter_set import FilterSet  # noqa: E402
from ._tracker import HeadTracker  # noqa: E402
from ._compensation import Compensation  # noqa: E402


def __getattr__(name):
    # `Convolver` needs the CUDA library; import lazily so that the filter-set and oracle-side
    # tooling can be used before the library is built, while any use of a convolver fails loudly.
    if name == "Convolver":
        from ._convolver import Convolver

        return Convolver
    raise AttributeError(name)
