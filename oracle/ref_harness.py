"""TEST INFRASTRUCTURE ONLY -- stub-import harness around the UNMODIFIED reference.

Imports ``/root/reference/ReTiSAR`` (read-only, exists only in the build container) -- or, on the
GPU box, the unmodified copy ``oracle/stage_ref.py`` staged into the git-ignored ``oracle/_ref`` --
with the third-party modules that are missing here replaced by mocks, so that the
reference's own ``OverlapSaveConvolver`` / ``AdjustableFdConvolver`` / ``AdjustableShConvolver``
(``ReTiSAR/_convolver.py``) and ``FilterSet*`` (``ReTiSAR/_filter_set.py``) can be executed on
seeded synthetic inputs.  ``oracle/make_golden.py`` uses it to write ``tests/golden/*.npz``;
those fixtures pin the numpy restatement in ``oracle/sh_chain.py``.

Nothing in ``retisar_b200/`` may import this module.

Recipe:
  * ``MagicMock`` for pythonosc, pysofaconventions, samplerate, soundfile, serial, pyfftw, jack,
    matplotlib (never exercised on the hot path);
  * a real mini ``sound_field_analysis`` module restating the five functions the path touches
    (package ``sound_field_analysis`` >=2021.2.4 per reference ``setup.py:51``; call sites
    ``_convolver.py:1095,1260`` and ``_filter_set.py:1123,1259,1263``);
  * ``config.IS_PYFFTW_MODE = False`` so the reference takes its ``np.fft`` branch
    (``_convolver.py:390,596,647``); ``np.math = math`` for ``_filter_set.py:616`` on numpy 2.
"""
import math
import os
import sys
import types
from collections import namedtuple
from unittest.mock import MagicMock

import numpy as np


def _find_reference_root():
    """The reference itself in the build container, else the unmodified copy that
    ``oracle/stage_ref.py`` staged into ``oracle/_ref`` (what the GPU box sees)."""
    env = os.environ.get("RTB_REFERENCE_ROOT")  # tests: force the staged copy
    if env:
        return env
    if os.path.isdir("/root/reference/ReTiSAR"):
        return "/root/reference"
    from oracle import stage_ref

    return stage_ref.staged_root() or "/root/reference"


REFERENCE_ROOT = _find_reference_root()


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ReTiSAR"))


def reference_is_original():
    """True when the package is read from /root/reference (build container), False for the staged copy."""
    return REFERENCE_ROOT == "/root/reference"


# --------------------------------------------------------------------------------------------
# mini `sound_field_analysis` (published algorithms of sound_field_analysis-py, restated)
# --------------------------------------------------------------------------------------------
def _mn_arrays(nMax):
    """sfa.sph.mnArrays: stacked orders m = [0,-1,0,1,-2,...] and degrees n = [0,1,1,1,2,...]."""
    m = np.concatenate([np.arange(-n, n + 1) for n in range(nMax + 1)])
    n = np.concatenate([np.full(2 * n + 1, n) for n in range(nMax + 1)])
    return m, n


def _reverse_mn_ids(nMax):
    """sfa.sph.reverseMnIds: index list mapping (n, m) -> (n, -m)."""
    ids = []
    for n in range(nMax + 1):
        first = n * n
        ids.extend(range(first + 2 * n, first - 1, -1))
    return ids


def _sph_harm_all(nMax, az, co):
    """sfa.sph.sph_harm_all: complex SH Y_n^m(az, colat) of size [points; (nMax+1)^2]."""
    from scipy.special import sph_harm_y

    m, n = _mn_arrays(nMax)
    mA, azA = np.meshgrid(m, np.asarray(az, dtype=np.float64).ravel())
    nA, coA = np.meshgrid(n, np.asarray(co, dtype=np.float64).ravel())
    return sph_harm_y(nA, mA, coA, azA)


def _spat_ft_rt(data, spherical_harmonic_weighted):
    """sfa.process.spatFT_RT: real-time spatial Fourier transform = plain matrix product."""
    return np.dot(spherical_harmonic_weighted, data)


ArrayConfiguration = namedtuple(
    "ArrayConfiguration",
    ["array_radius", "array_type", "transducer_type", "scatter_radius", "dual_radius"],
)
SphericalGrid = namedtuple("SphericalGrid", ["azimuth", "colatitude", "radius", "weight"])


def _build_sfa_module():
    sfa = types.ModuleType("sound_field_analysis")
    for sub in ("process", "sph", "io", "gen", "utils", "plot"):
        setattr(sfa, sub, types.ModuleType(f"sound_field_analysis.{sub}"))
    sfa.process.spatFT_RT = _spat_ft_rt
    sfa.sph.mnArrays = _mn_arrays
    sfa.sph.reverseMnIds = _reverse_mn_ids
    sfa.sph.sph_harm_all = _sph_harm_all
    sfa.io.ArrayConfiguration = ArrayConfiguration
    sfa.io.SphericalGrid = SphericalGrid
    return sfa


_REF = None


def import_reference():
    """Import and return the reference package (cached)."""
    global _REF
    if _REF is not None:
        return _REF
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT} (build container only)")

    for name in (
        "pythonosc", "pythonosc.dispatcher", "pythonosc.osc_server", "pythonosc.udp_client",
        "pysofaconventions", "samplerate", "soundfile", "serial", "serial.tools",
        "serial.tools.list_ports", "pyfftw", "pyfftw.builders",
        "matplotlib", "matplotlib.pyplot", "matplotlib.ticker", "natsort",
    ):
        sys.modules.setdefault(name, MagicMock())
    from oracle import fake_jack

    sys.modules.setdefault("jack", fake_jack.build_module())
    sfa = _build_sfa_module()
    sys.modules["sound_field_analysis"] = sfa
    for sub in ("process", "sph", "io", "gen", "utils", "plot"):
        sys.modules[f"sound_field_analysis.{sub}"] = getattr(sfa, sub)
    if not hasattr(np, "math"):
        np.math = math  # `_filter_set.py:616` uses the numpy-1 alias

    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import ReTiSAR  # noqa
    finally:
        sys.path.remove(REFERENCE_ROOT)
    ReTiSAR.config.IS_PYFFTW_MODE = False
    ReTiSAR.config.IS_DEBUG_MODE = False
    ReTiSAR.config.IS_RUNNING.set()
    _REF = ReTiSAR
    return _REF


# --------------------------------------------------------------------------------------------
# builders of reference objects from synthetic ndarrays
# --------------------------------------------------------------------------------------------
def make_tracker():
    """`multiprocessing.Array('f', 6)` like `HeadTracker` (`_tracker.py:73-81,156-164`)."""
    ref = import_reference()
    return ref.mp_context.Array("f", 6)


def make_miro_filter_set(irs_td, azimuth, colatitude, weight, sh_max_order, is_hrir,
                         block_length, fs=48000, enforce_pinv=False, array_radius=0.042):
    """Reference `FilterSetMiro` filled from arrays instead of a `.mat` file.

    Mirrors what `FilterSetMiro._load` (`_filter_set.py:996-1074`) and `FilterSet.load`
    (`:254-356`) leave behind, then runs the reference's own `_zero_pad` and dirac generation.
    """
    ref = import_reference()
    fsm = ref._filter_set.FilterSetMiro(
        file_name="SYNTHETIC", is_hrir=is_hrir, sh_max_order=sh_max_order,
        sh_is_enforce_pinv=enforce_pinv,
    )
    dtype = irs_td.dtype
    fsm._irs_td = np.ascontiguousarray(irs_td)
    fsm._fs = fs
    fsm._irs_grid = SphericalGrid(
        np.asarray(azimuth, dtype=dtype), np.asarray(colatitude, dtype=dtype),
        np.ones_like(np.asarray(azimuth, dtype=dtype)),
        None if weight is None else np.asarray(weight, dtype=dtype),
    )
    if is_hrir:
        fsm._points_azim_deg = np.rad2deg(fsm._irs_grid.azimuth)
        fsm._points_elev_deg = 90 - np.rad2deg(fsm._irs_grid.colatitude)
    else:
        fsm._arir_config = ArrayConfiguration(
            np.asarray(array_radius, dtype=dtype), "rigid", "omni",
            np.asarray(array_radius, dtype=dtype), None,
        )
    fsm._irs_orig_shape = fsm._irs_td.shape
    fsm._zero_pad(block_length=block_length)
    fsm._dirac_td = np.zeros(fsm._irs_td.shape[-2:], dtype=fsm._irs_td.dtype)
    fsm._dirac_td[:, 0] = 1.0
    return fsm


def make_multichannel_filter_set(irs, block_length, is_hpcf=False, single_precision=True):
    """Reference `FilterSetMultiChannel` from an ndarray [channels; taps] (`_filter_set.py:786`)."""
    ref = import_reference()
    ftype = "HPCF_FIR" if is_hpcf else "FIR_MULTICHANNEL"
    fs = ref.FilterSet.create_instance_by_type(np.ascontiguousarray(irs), ftype)
    fs.load(block_length=block_length, is_single_precision=single_precision,
            is_prevent_logging=True)
    return fs


def make_sh_convolver(hrir_set, array_set, block_length, tracker, comp_nm=None,
                      source_positions=((0, 0),)):
    """Reference `AdjustableShConvolver` prepared without the plotting / `sfa.gen` parts.

    Runs the non-plotting lines of `prepare_sh_processing` (`_convolver.py:1094-1101,1113-1116`)
    and `filter.calculate_filter_blocks_nm()` (`:1189`); the compensation spectra `comp_nm`
    ([K;1;F] complex, `:1206`) are supplied by the caller (identity if None).
    """
    ref = import_reference()
    conv = ref._convolver.AdjustableShConvolver(hrir_set, block_length, list(source_positions),
                                                tracker)
    cfg = array_set.get_sh_configuration()
    sfa = sys.modules["sound_field_analysis"]
    conv._sh_m = cfg.sh_m.copy()
    conv._sh_m_rev_id = sfa.sph.reverseMnIds(hrir_set._sh_max_order)
    conv._sh_cur_order = hrir_set._sh_max_order
    conv._sh_bases_weighted = cfg.sh_bases_weighted.copy()
    conv._last_sh_azim_nm = np.zeros((conv._sh_bases_weighted.shape[0], 1, 1),
                                     dtype=conv._sh_bases_weighted.dtype)
    hrir_set.calculate_filter_blocks_nm()
    if comp_nm is not None:
        hrir_set._irs_blocks_nm *= comp_nm
    conv._input_block_td = np.zeros((conv._sh_bases_weighted.shape[-1], block_length * 2),
                                    dtype=hrir_set.get_dirac_td().dtype)
    return conv


def make_ssr_filter_set(irs_td, block_length, is_hrir=True, fs=48000):
    """Reference `FilterSetSsr` filled from an array [directions; 2; taps] instead of the 720-channel
    WAV that `FilterSetSsr._load` reads (`_filter_set.py:852-881`)."""
    ref = import_reference()
    fset = ref._filter_set.FilterSetSsr(file_name="SYNTHETIC", is_hrir=is_hrir)
    fset._irs_td = np.ascontiguousarray(irs_td)
    fset._fs = fs
    fset._points_azim_deg = np.linspace(0, 360, irs_td.shape[0], endpoint=False, dtype=np.int16)
    fset._points_elev_deg = np.zeros_like(fset._points_azim_deg)
    fset._irs_orig_shape = fset._irs_td.shape
    fset._zero_pad(block_length=block_length)
    fset._dirac_td = np.zeros(fset._irs_td.shape[-2:], dtype=fset._irs_td.dtype)
    fset._dirac_td[:, 0] = 1.0
    return fset


def inject_compensation_backend():
    """Give the stub `sound_field_analysis` the design functions `ReTiSAR/_compensation.py` calls
    (`gen.radial_filter_fullspec` :586, `process.rfi` :595, `gen.spherical_head_filter_spec` :659,
    `gen.delay_fd` :670, `gen.tapering_window` :713, `utils.zero_pad_fd` :505), bound to the
    restatements in `retisar_b200/_compensation.py` under the third-party package's call
    signatures.  With them the reference's OWN `Compensation.generate_by_type` runs unmodified, so
    its bookkeeping -- type parsing, once-only application, filter-length budget, per-degree
    repetition, the (-1)^m factor of the radial filters, SDS / EDS masks, zero padding -- produces
    goldens for this repository's `Compensation`.  The design VALUES stay restated (the package is
    absent): those remain parity-unpinned."""
    import_reference()
    from retisar_b200 import _compensation as own

    sfa = sys.modules["sound_field_analysis"]
    sfa.gen.radial_filter_fullspec = lambda max_order, NFFT, fs, array_configuration, amp_maxdB=40: \
        own.radial_filter_fullspec(max_order, NFFT, fs, array_configuration, amp_maxdB)
    sfa.process.rfi = lambda FIR, kernelSize=512: own.rfi(FIR, kernelSize)
    sfa.gen.spherical_head_filter_spec = lambda max_order, NFFT, fs, radius, is_tapering=False: \
        own.spherical_head_filter_spec(max_order, NFFT, fs, radius, is_tapering)
    sfa.gen.delay_fd = lambda target_length_fd, delay_samples: own.delay_fd(target_length_fd, delay_samples)
    sfa.gen.tapering_window = lambda max_order: own.tapering_window(max_order)
    sfa.utils.zero_pad_fd = lambda data_fd, target_length_td: own.zero_pad_fd(data_fd, target_length_td)

