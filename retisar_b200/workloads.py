"""Synthetic render-chain configurations of BASELINE.json, shared by
``bench.py`` and the tests.  All data is seeded and generated in memory."""
import numpy as np

from . import synthetic as syn
from ._filter_set import FilterSetMiro

FS = 48000

# name -> (array grid, SH order, block length, HRIR taps, description)
CONFIGS = {
    "c1": dict(grid="em32", order=4, block=256, hrir_taps=128,
               desc="Eigenmike em32 (32 ch), SH order 4, 256-sample block, 128-tap KU100-shaped HRIR, "
                    "ONE stream (BASELINE config 1: the reference's own CPU-runnable case)"),
    "b4096": dict(grid="em32", order=4, block=4096, hrir_taps=128,
                  desc="Eigenmike em32 (32 ch), SH order 4, 4096-sample block (the reference's default "
                       "block length, config.py:9), 128-tap HRIR"),
    "c2": dict(grid="em32", order=4, block=256, hrir_taps=128,
               desc="Eigenmike em32 (32 ch), SH order 4, 256-sample block, 128-tap KU100-shaped HRIR"),
    "c3": dict(grid="le110", order=8, block=256, hrir_taps=4096,
               desc="Lebedev 110-point array, SH order 8, 256-sample block, 4096-tap HRIR as 16 "
                    "uniform partitions (extension a-X)"),
    "c5": dict(grid="le110", order=8, block=256, hrir_taps=128,
               desc="Lebedev 110-point array, SH order 8, 256-sample block, 128-tap HRIR"),
    "c4": dict(grid="rnd434", order=12, block=64, hrir_taps=64,
               desc="434-point array, SH order 12 (169 coeffs), 64-sample low-latency block"),
}


def _grid(name):
    if name == "rnd434":  # no 434-point grid ships with the reference: pseudo-random + pinv
        az, co = syn.random_sphere_grid(434, 434)
        return az, co, None
    return syn.load_array_grid(name)


def sh_tables(name, hrir_dirs=2702, seed=7):
    """Filter sets and tables of one configuration.

    Returns dict(array_set, hrir_set, sh_config, comp_nm, B, order, C, K, F): the HRIR set is
    "KU100-shaped" (2702 directions, the reference's default set ``config.py:48``) on a seeded
    pseudo-random grid with pinv weights (SOFA sets carry none); compensation = identity.
    """
    c = CONFIGS[name]
    B, order = c["block"], c["order"]
    az, co, w = _grid(c["grid"])
    array_set = FilterSetMiro.from_arrays(np.zeros((1, len(az), B), np.float32), az, co, w, order,
                                          is_hrir=False)
    array_set.load(B, True, is_prevent_logging=True)
    haz, hco = syn.random_sphere_grid(hrir_dirs, seed)
    hrir_set = FilterSetMiro.from_arrays(
        syn.decaying_noise_irs((hrir_dirs, 2, c["hrir_taps"]), seed + 1), haz, hco, None, order,
        is_hrir=True)
    hrir_set.allow_partitions = c["hrir_taps"] > B  # extension a-X (the reference refuses P > 1)
    hrir_set.load(B, True, is_prevent_logging=True)
    K, F = (order + 1) ** 2, B + 1
    return dict(array_set=array_set, hrir_set=hrir_set, sh_config=array_set.get_sh_configuration(),
                comp_nm=np.ones((K, 1, F), np.complex64), B=B, order=order, C=len(az), K=K, F=F,
                desc=c["desc"])


def make_renderer(tables, n_streams, device=None):
    """``AdjustableShConvolver`` for `n_streams` batched streams, prepared and ready to run."""
    from ._convolver import Convolver

    conv = Convolver.create_instance_by_filter_set(
        tables["hrir_set"], tables["B"], [(0, 0)], None, n_streams=n_streams, device=device)
    conv.prepare_sh_processing(tables["sh_config"], 0, None, comp_nm=tables["comp_nm"])
    return conv


def chain_bytes_per_stream_block(C, B, E=2):
    """Compulsory HBM bytes of the whole chain per stream and block:
    read current block + read previous-block overlap state + write output + yaw."""
    return 4 * B * (2 * C + E) + 12


def chain_flops_per_stream_block(C, K, B, E=2):
    """fp32 flops of the reference formulation per stream and block."""
    F, N2 = B + 1, 2 * B
    fft = 2.5 * N2 * np.log2(N2)
    return C * fft + 8 * K * C * F + 6 * K * E * F + 2 * 8 * K * E * F + 2 * E * fft + 3 * E * B


def make_c4_chain(tables, n_streams, arir_taps=4096, hpcf_taps=1024, device=None, seed=40):
    """The three convolvers of BASELINE config c4 in the reference's production topology
    (``JackRenderer`` chain pre-renderer -> renderer -> HPCF, ``_run.py:setup_*``): a mono source
    through a 434-channel ARIR (``OverlapSaveConvolver``, mono input broadcast,
    ``_convolver.py:588-591``), the SH renderer, and a 2-channel headphone compensation filter.

    Returns (arir, renderer, hpcf, arir_irs [C; taps], hpcf_irs [2; taps]); SURVEY.md section 8d:
    seeded decaying noise, 4096-tap ARIR (P = 64 at B = 64), 1024-tap HPCF (P = 16).
    """
    from ._convolver import Convolver
    from ._filter_set import FilterSet

    B, C = tables["B"], tables["C"]
    arir_irs = syn.decaying_noise_irs((C, arir_taps), seed)
    hpcf_irs = syn.decaying_noise_irs((2, hpcf_taps), seed + 1, scale=0.3)
    convs = []
    for irs in (arir_irs, hpcf_irs):
        fs = FilterSet.create_instance_by_type(irs, "FIR_MULTICHANNEL")
        fs.load(block_length=B, is_single_precision=True, is_prevent_logging=True)
        convs.append(Convolver.create_instance_by_filter_set(fs, B, n_streams=n_streams, device=device))
    return convs[0], make_renderer(tables, n_streams, device=device), convs[1], arir_irs, hpcf_irs
