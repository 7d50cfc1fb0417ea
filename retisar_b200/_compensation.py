"""SH-domain compensation spectra (reference ``ReTiSAR/_compensation.py``) -- set-up time only.

The render path consumes a table ``comp_nm [K; 1; F]`` that is multiplied once into the HRIR SH
coefficients (``_convolver.py:1195-1206``).  The filter designs themselves (modal radial filters,
spherical head filter, SH tapering; ``_compensation.py:556-733``) are a later row of the scope
table; until they exist callers supply ``comp_nm`` to
``AdjustableShConvolver.prepare_sh_processing``.
"""
from enum import Enum


class Compensation(object):
    class Type(Enum):
        """Identifiers as in ``_compensation.py:45-180`` (long names are aliases)."""

        SPF = "SUBSONIC_PRE_FILTER"
        SUBSONIC_PRE_FILTER = SPF
        MRF = "MODAL_RADIAL_FILTER"
        MODAL_RADIAL_FILTER = MRF
        SHF = "SPHERICAL_HEAD_FILTER"
        SPHERICAL_HEAD_FILTER = SHF
        SHT = "SPHERICAL_HARMONICS_TAPERING"
        SPHERICAL_HARMONICS_TAPERING = SHT
        SDS = "SECTORIAL_DEGREE_SELECTION"
        SECTORIAL_DEGREE_SELECTION = SDS
        EDS = "EQUATORIAL_DEGREE_SELECTION"
        EQUATORIAL_DEGREE_SELECTION = EDS

    @staticmethod
    def generate_by_type(compensation_types, filter_set, arir_config, amp_limit_db, nfft,
                         nfft_padded, sh_cur_order=None, logger=None):
        raise NotImplementedError(
            "compensation filter design is not built yet; pass `comp_nm` to "
            "`prepare_sh_processing` (identity: numpy.ones((K, 1, F)))")
