"""TEST INFRASTRUCTURE ONLY -- generate ``tests/golden/*.npz`` by running the UNMODIFIED reference.

Run in the build container (needs ``/root/reference``):  ``python -m oracle.make_golden``.
Every fixture stores the inputs, the tables the reference derived, and the outputs the
reference's own classes produced block by block, so the oracle and the CUDA path can be checked
without the reference being present (it does not exist on the GPU box).

Also extracts the two measured array geometries used by the benchmarks into
``retisar_b200/data/array_grids.npz``.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import ref_harness as rh  # noqa: E402
from retisar_b200 import synthetic as syn  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(ROOT, "retisar_b200", "data")


def extract_grids():
    import scipy.io as sio

    out = {}
    for name, fn in (("em32", "RT_calib_EM32ch_struct.mat"),
                     ("le110", "DRIR_sim_LE110_PW_struct.mat")):
        m = sio.loadmat(os.path.join(rh.REFERENCE_ROOT, "res", "ARIR", fn))
        out[f"{name}_azimuth"] = m["azimuth"].ravel().astype(np.float64)
        out[f"{name}_colatitude"] = m["colatitude"].ravel().astype(np.float64)
        out[f"{name}_weight"] = m["quadWeight"].ravel().astype(np.float64)
        out[f"{name}_radius"] = m["radius"].ravel().astype(np.float64)
    os.makedirs(DATA, exist_ok=True)
    np.savez_compressed(os.path.join(DATA, "array_grids.npz"), **out)


def _array_and_hrir(grid, order, B, hrir_dirs, hrir_taps, seed, enforce_pinv=False, array_radius=0.042):
    """Reference FilterSetMiro pair: (array set for Y_w, HRIR set for H_nm)."""
    az, co, w = grid
    arr = rh.make_miro_filter_set(
        np.zeros((1, len(az), B), dtype=np.float32), az, co, w, order, is_hrir=False,
        block_length=B, enforce_pinv=enforce_pinv, array_radius=array_radius)
    haz, hco = syn.random_sphere_grid(hrir_dirs, seed + 1)
    hrir = rh.make_miro_filter_set(
        syn.decaying_noise_irs((hrir_dirs, 2, hrir_taps), seed + 2), haz, hco, None, order,
        is_hrir=True, block_length=B)
    hrir.calculate_filter_blocks_fd(B)
    return arr, hrir


def golden_sh_chain(name, grid, order, B, T, seed, events=None, hrir_dirs=200, hrir_taps=None,
                    yaws=None, comp=False, random_yw=False, enforce_pinv=False,
                    source_azim=0.0, is_hrir_flag=True):
    """AdjustableShConvolver.filter_block (`_convolver.py:1231-1316`) over T blocks."""
    ref = rh.import_reference()
    events = events or {}
    arr, hrir = _array_and_hrir(grid, order, B, hrir_dirs, hrir_taps or min(B, 128), seed,
                                enforce_pinv)
    K = (order + 1) ** 2
    F = B + 1
    comp_nm = None
    if comp:
        rng = np.random.default_rng(seed + 3)
        comp_nm = (rng.standard_normal((K, 1, F)) + 1j * rng.standard_normal((K, 1, F))
                   ).astype(np.complex64)
    tracker = rh.make_tracker()
    conv = rh.make_sh_convolver(hrir, arr, B, tracker, comp_nm=comp_nm,
                                source_positions=((source_azim, 0),))
    hrir._is_hrir = is_hrir_flag
    if random_yw:
        rng = np.random.default_rng(seed + 4)
        conv._sh_bases_weighted = (
            (rng.standard_normal(conv._sh_bases_weighted.shape)
             + 1j * rng.standard_normal(conv._sh_bases_weighted.shape)) * 0.2
        ).astype(np.complex64)
    C = conv._sh_bases_weighted.shape[1]
    x = syn.noise_blocks(T, C, B, seed)
    if yaws is None:
        yaws = syn.yaw_walk(T, seed + 5)
    yaws = np.asarray(yaws, dtype=np.float32)
    outs = np.zeros((T, 2, B), dtype=np.float32)
    ev_codes = np.zeros((T, 2), dtype=np.int32)  # (kind, value) applied BEFORE block t
    EV = {"order": 1, "crossfade": 2, "passthrough": 3, "clear": 4}
    for t in range(T):
        if t in events:
            kind, val = events[t]
            ev_codes[t] = (EV[kind], int(val))
            if kind == "order":
                conv._sh_cur_order = int(val)  # `_convolver.py:1160`
            elif kind == "crossfade":
                conv.set_crossfade(bool(val))
            elif kind == "passthrough":
                conv.set_passthrough(bool(val))
            elif kind == "clear":
                conv._clear_buffers()
        tracker[ref.HeadTracker.DataIndex.AZIM] = float(yaws[t])
        y = conv.filter_block(x[t].copy())
        outs[t] = y
    np.savez_compressed(
        os.path.join(GOLD, f"{name}.npz"),
        kind="sh_chain", order=order, B=B, sh_m=conv._sh_m, rev=np.asarray(conv._sh_m_rev_id),
        Yw=conv._sh_bases_weighted, Hnm=hrir._irs_blocks_nm, x=x, yaw=yaws, y=outs,
        events=ev_codes, source_azim=np.float32(source_azim), is_hrir=is_hrir_flag,
        w_in=conv._window_in_td, w_out=conv._window_out_td,
    )
    print(name, "max|y|", np.abs(outs).max())


def golden_ols(name, irs, B, T, seed, mono_in=False, events=None, is_hpcf=False):
    """OverlapSaveConvolver.filter_block (`_convolver.py:524-571`) over T blocks."""
    events = events or {}
    ref = rh.import_reference()
    fset = rh.make_multichannel_filter_set(irs, B, is_hpcf=is_hpcf)
    conv = ref._convolver.OverlapSaveConvolver(fset, B)
    n_in = 1 if mono_in else irs.shape[0]
    x = syn.noise_blocks(T, n_in, B, seed)
    outs = np.zeros((T, irs.shape[0], B), dtype=np.float32)
    ev_codes = np.zeros((T, 2), dtype=np.int32)
    for t in range(T):
        if t in events:
            kind, val = events[t]
            ev_codes[t] = ({"passthrough": 3, "clear": 4}[kind], int(val))
            if kind == "passthrough":
                conv.set_passthrough(bool(val))
            else:
                conv._clear_buffers()
        outs[t] = conv.filter_block(x[t].copy())
    np.savez_compressed(
        os.path.join(GOLD, f"{name}.npz"), kind="ols", B=B, irs=irs, x=x, y=outs,
        Hfd=fset.get_filter_blocks_fd(), events=ev_codes, irs_td_padded=fset._irs_td,
    )
    print(name, "max|y|", np.abs(outs).max())


def golden_fd(name, n_dirs, taps, B, T, seed, sources, is_hrir=True, events=None):
    """AdjustableFdConvolver.filter_block (`_convolver.py:780-866`) over T blocks."""
    events = events or {}
    ref = rh.import_reference()
    irs = syn.decaying_noise_irs((n_dirs, 2, taps), seed)
    fset = rh.make_ssr_filter_set(irs, B, is_hrir=is_hrir)
    tracker = rh.make_tracker()
    conv = ref._convolver.AdjustableFdConvolver(fset, B, list(sources), tracker)
    x = syn.noise_blocks(T, len(sources), B, seed + 1)
    yaws = syn.yaw_walk(T, seed + 2, step=40.0)
    outs = np.zeros((T, 2, B), dtype=np.float32)
    ev_codes = np.zeros((T, 2), dtype=np.int32)
    EV = {"crossfade": 2, "passthrough": 3, "clear": 4}
    for t in range(T):
        if t in events:
            kind, val = events[t]
            ev_codes[t] = (EV[kind], int(val))
            if kind == "crossfade":
                conv.set_crossfade(bool(val))
            elif kind == "passthrough":
                conv.set_passthrough(bool(val))
            else:
                conv._clear_buffers()
        tracker[ref.HeadTracker.DataIndex.AZIM] = float(yaws[t])
        outs[t] = conv.filter_block(x[t].copy())
    np.savez_compressed(
        os.path.join(GOLD, f"{name}.npz"), kind="fd", B=B, x=x, yaw=yaws, y=outs,
        # the 360-direction filter set is regenerated from its seed (syn.decaying_noise_irs);
        # the checksum guards against a drifting random stream
        irs_shape=np.asarray(irs.shape), irs_seed=seed, irs_sum=np.float64(np.abs(irs.astype(np.float64)).sum()),
        sources=np.asarray(sources, dtype=np.float32), is_hrir=is_hrir, events=ev_codes,
    )
    print(name, "max|y|", np.abs(outs).max())


def golden_tables(name, grid, order, B, seed, enforce_pinv=False):
    """get_sh_configuration / calculate_filter_blocks_fd / _nm (`_filter_set.py:625-664,1093-1280`)."""
    arr, hrir = _array_and_hrir(grid, order, B, 24, min(B, 100), seed, enforce_pinv)
    cfg = arr.get_sh_configuration()
    hrir.calculate_filter_blocks_nm()
    hcfg = hrir.get_sh_configuration()
    np.savez_compressed(
        os.path.join(GOLD, f"{name}.npz"), kind="tables", order=order, B=B,
        az=grid[0], co=grid[1], w=np.zeros(0) if grid[2] is None else grid[2],
        enforce_pinv=enforce_pinv, sh_m=cfg.sh_m, Yw=cfg.sh_bases_weighted,
        hrir_az=hrir._irs_grid.azimuth, hrir_co=hrir._irs_grid.colatitude,
        hrir_td=hrir._irs_td[..., :min(B, 100)], hrir_td_padded_shape=np.asarray(hrir._irs_td.shape),
        hrir_blocks_fd=hrir._irs_blocks_fd, hrir_Yw=hcfg.sh_bases_weighted,
        hrir_blocks_nm=hrir._irs_blocks_nm,
    )
    print(name, cfg.sh_bases_weighted.shape)


def golden_yaw(name):
    """_calculate_individual_directions + phase build (`_convolver.py:947-975,1278-1283`)."""
    ref = rh.import_reference()
    rng = np.random.default_rng(77)
    yaws = np.concatenate([
        syn.YAW_EDGE_CASES, rng.uniform(-1000, 1000, 250).astype(np.float32),
        np.arange(0, 360, 0.125, dtype=np.float32)[::29],
        np.array([65000.0, -65000.0, 1e-3, -1e-3, 359.875, 359.9375, 360.0, -360.0, 719.9],
                 dtype=np.float32),
    ]).astype(np.float32)
    m = np.arange(-12, 13).astype(np.int16)  # every order m an SH order <= 12 can produce
    out = {}
    for tag, src, is_hrir in (("hrir", 0.0, True), ("src", 33.3, True), ("brir", 0.0, False)):
        arr, hrir = _array_and_hrir(syn.random_sphere_grid(8, 1) + (None,), 1, 32, 8, 16, 5)
        tracker = rh.make_tracker()
        conv = rh.make_sh_convolver(hrir, arr, 32, tracker, source_positions=((src, 0),))
        hrir._is_hrir = is_hrir
        az = np.zeros(len(yaws), dtype=np.float16)
        ph = np.zeros((len(yaws), len(m)), dtype=np.complex64)
        with np.errstate(all="ignore"):
            for i, y in enumerate(yaws):
                tracker[ref.HeadTracker.DataIndex.AZIM] = float(y)
                a, _ = conv._calculate_individual_directions()
                az[i] = a[0]
                ph[i] = np.exp(np.complex64(-1j) * m * np.deg2rad(a))
        out[f"az_{tag}"] = az
        out[f"ph_{tag}"] = ph
    np.savez_compressed(os.path.join(GOLD, f"{name}.npz"), kind="yaw", yaw=yaws, sh_m=m,
                        src_azim=np.float32(33.3), **out)
    print(name, len(yaws))


def golden_delay(name, fs=48000, B=64, channels=3, max_sec=0.01, T=30, seed=401):
    """DelayBuffer (`_convolver.py:11-141`): ring of blocks, delay changes while running, the
    maximum-delay wrap (delay == ring size returns the block just written) and `None` input."""
    ref = rh.import_reference()
    buf = ref._convolver.DelayBuffer(sample_rate=fs, block_length=B, delay_ms=2.0)
    buf.init(max_length_sec=max_sec, channel_count=channels)
    x = syn.noise_blocks(T, channels, B, seed)
    events = {5: 0.0, 8: 4.0, 14: 10.0, 18: 1000.0, 24: 1.4}  # block -> set_delay(ms)
    ev = np.full(T, -1.0, dtype=np.float64)
    realized = np.full(T, -1.0, dtype=np.float64)
    outs = np.zeros((T, channels, B), dtype=np.float32)
    for t in range(T):
        if t in events:
            ev[t] = events[t]
            realized[t] = buf.set_delay(events[t])
        outs[t] = np.array(buf.process_block(x[t].copy()))
    assert buf.process_block(None) is None
    np.savez_compressed(os.path.join(GOLD, f"{name}.npz"), kind="delay", fs=fs, B=B, max_sec=max_sec,
                        start_delay_ms=2.0, x=x, y=outs, set_delay_ms=ev, realized_ms=realized,
                        ring_blocks=buf._buffer.shape[0])
    print(name, "ring blocks", buf._buffer.shape[0])


def golden_levels(name, seed=402):
    """`tools.calculate_rms` / `calculate_peak` (`tools.py:1117-1169`) on output blocks scaled like
    `_process_deliver` (`_jack_client.py:795`), incl. an all-zero channel and a clipping one."""
    ref = rh.import_reference()
    x = syn.noise_blocks(6, 4, 256, seed, scale=0.3)
    x[2, 1] = 0.0
    x[4, 3] *= 9.0
    volume, rel = 10 ** (-3.0 / 20.0), np.array([[1.0], [0.5], [2.0], [1.0]])
    y = x.copy()
    rms, peak = [], []
    for t in range(x.shape[0]):
        y[t] *= volume * rel
        rms.append(ref.tools.calculate_rms(y[t], is_level=True))
        peak.append(ref.tools.calculate_peak(y[t], is_level=True))
    np.savez_compressed(os.path.join(GOLD, f"{name}.npz"), kind="levels", x=x, y=y, volume=volume,
                        relative=rel, rms=np.stack(rms), peak=np.stack(peak))
    print(name, "rms", rms[0])


def golden_compensation(name="compensation_em32_le110"):
    """The reference's own `Compensation.generate_by_type` (`_compensation.py:182-505`) with the
    third-party design functions bound to the restatements (`ref_harness.inject_compensation_backend`):
    pins the bookkeeping of `retisar_b200/_compensation.py` -- type lists and "+"-joined strings,
    once-only application, filter-length budget, per-degree repetition, the (-1)^m factor of the
    radial filters (`:617-627`), SHF with / without tapering, SDS / EDS masks (`:734-833`), zero
    padding to the block spectrum length."""
    ref = rh.import_reference()
    rh.inject_compensation_backend()
    C = ref.Compensation
    em32 = syn.load_array_grid("em32")
    le110 = syn.load_array_grid("le110")
    cases = [
        ("sht_n4_b256", em32, 4, 256, 128, ["SHT"], 18, 0.042),
        ("sds_n4_b256", em32, 4, 256, 128, ["SDS"], 18, 0.042),
        ("eds_n8_b128", le110, 8, 128, 64, ["EDS"], 18, 0.0875),
        ("mrf_n4_b256", em32, 4, 256, 128, ["MRF"], 18, 0.042),
        ("shf_mrf_n4_b256", em32, 4, 256, 128, ["SHF", "MRF"], 18, 0.042),
        ("sht_shf_mrf_n8_b256", le110, 8, 256, 100, ["SHT+SHF", "MRF"], 40, 0.0875),
        ("sds_mrf_mrf_n3_b512", em32, 3, 512, 50, ["SDS", "MRF", "MRF"], 6, 0.042),
    ]
    out = {"names": np.array([c[0] for c in cases])}
    for cname, grid, order, B, taps, types, limit_db, radius in cases:
        arr, hrir = _array_and_hrir(grid, order, B, 60, taps, 500, array_radius=radius)
        hrir.calculate_filter_blocks_nm()
        cfg = arr.get_sh_configuration()
        C.reset_config(is_plot=False)
        comp = C.generate_by_type(types, filter_set=hrir, arir_config=cfg.arir_config,
                                  amp_limit_db=limit_db, nfft=None, nfft_padded=2 * B)
        out[cname] = np.asarray(comp)
        out[cname + "_meta"] = np.array([order, B, taps, limit_db, radius], dtype=np.float64)
        out[cname + "_types"] = np.array(types)
        print(name, cname, comp.shape, comp.dtype, float(np.abs(comp).max()))
    np.savez_compressed(os.path.join(GOLD, f"{name}.npz"), **out)


def golden_production_chain(name="chain_em32_n4_b64"):
    """The reference's production topology run with its own classes end to end
    (`_run.py:240-338`): mono source -> ARIR pre-renderer (`OverlapSaveConvolver`, mono input
    broadcast, `_convolver.py:588-591`) -> `AdjustableShConvolver` -> HPCF (`OverlapSaveConvolver`,
    2 channels).  Pins the folded forms of this repository (`set_pre_renderer`, `set_hpcf`) directly
    to the unmodified reference."""
    ref = rh.import_reference()
    em32 = syn.load_array_grid("em32")
    order, B, T = 4, 64, 22
    arr, hrir = _array_and_hrir(em32, order, B, 120, 64, 700)
    C = len(em32[0])
    arir_irs = syn.decaying_noise_irs((C, 200), 701)
    hpcf_irs = syn.decaying_noise_irs((2, 300), 702, scale=0.3)
    tracker = rh.make_tracker()
    rend = rh.make_sh_convolver(hrir, arr, B, tracker)
    arir = ref._convolver.OverlapSaveConvolver(rh.make_multichannel_filter_set(arir_irs, B), B)
    hpcf = ref._convolver.OverlapSaveConvolver(rh.make_multichannel_filter_set(hpcf_irs, B, is_hpcf=False), B)
    x = syn.noise_blocks(T, 1, B, 703, scale=0.5)
    yaws = syn.yaw_walk(T, 704, step=20.0)
    mid = np.zeros((T, 2, B), dtype=np.float32)
    out = np.zeros((T, 2, B), dtype=np.float32)
    for t in range(T):
        if t == 15:  # the pipeline is paused: every client clears its buffers (`_convolver.py:383-386`)
            for c in (arir, rend, hpcf):
                c._clear_buffers()
        tracker[ref.HeadTracker.DataIndex.AZIM] = float(yaws[t])
        a = arir.filter_block(x[t].copy())
        r = rend.filter_block(np.array(a, copy=True))
        mid[t] = r
        out[t] = hpcf.filter_block(np.array(r, copy=True))
    cfg = arr.get_sh_configuration()
    np.savez_compressed(
        os.path.join(GOLD, f"{name}.npz"), kind="chain", order=order, B=B, x=x, yaw=yaws, y_renderer=mid, y=out,
        arir_irs=arir_irs, hpcf_irs=hpcf_irs, sh_m=cfg.sh_m, Yw=cfg.sh_bases_weighted, Hnm=hrir._irs_blocks_nm,
        rev=np.asarray(rend._sh_m_rev_id), w_in=rend._window_in_td, w_out=rend._window_out_td, clear_at=15)
    print(name, "max|y|", np.abs(out).max(), "max|renderer|", np.abs(mid).max())


def main():
    os.makedirs(GOLD, exist_ok=True)
    extract_grids()
    em32 = syn.load_array_grid("em32")
    le110 = syn.load_array_grid("le110")
    rnd50 = syn.random_sphere_grid(50, 50) + (None,)

    golden_yaw("yaw_chain")
    golden_tables("tables_em32_n4", em32, 4, 256, 11)
    golden_tables("tables_le110_n8_pinv", le110, 8, 128, 12, enforce_pinv=True)
    golden_tables("tables_rnd50_n3", rnd50, 3, 64, 13)

    # c1: em32 order 4, B=256, yaw edge cases (wraps), crossfade on
    golden_sh_chain("sh_em32_n4_b256", em32, 4, 256, 8, 101, yaws=syn.YAW_EDGE_CASES[:8])
    # control events on a small block: order lowering/raising, crossfade toggles, passthrough
    golden_sh_chain(
        "sh_em32_n4_b64_events", em32, 4, 64, 24, 102, comp=True,
        events={4: ("order", 2), 7: ("order", 4), 9: ("crossfade", 0), 12: ("crossfade", 1),
                14: ("passthrough", 1), 16: ("passthrough", 0), 18: ("order", 0),
                19: ("order", 3), 21: ("clear", 0)})
    # c3 geometry (Lebedev 110, order 8), shorter block to keep the fixture small
    golden_sh_chain("sh_le110_n8_b128", le110, 8, 128, 4, 103, comp=True)
    # pinv grid (weightless), order 3
    golden_sh_chain("sh_rnd50_n3_b32", rnd50, 3, 32, 12, 104, comp=True)
    # non-symmetric Y_w (general complex analysis matrix), BRIR direction flag, source offset
    golden_sh_chain("sh_em32_n2_b64_general", em32, 2, 64, 10, 105, random_yw=True, comp=True,
                    source_azim=33.3, is_hrir_flag=False)
    # order 12 on a 434-point pseudo-random grid (c4 geometry), 64-sample block
    rnd434 = syn.random_sphere_grid(434, 434) + (None,)
    golden_sh_chain("sh_rnd434_n12_b64", rnd434, 12, 64, 3, 106, hrir_dirs=400)

    # HRIR / BRIR switching renderer (SSR-style sets)
    golden_fd("fd_hrir_1src_b64", 360, 150, 64, 20, 301, [(0, 0)],
              events={7: ("crossfade", 0), 10: ("crossfade", 1), 13: ("passthrough", 1),
                      15: ("passthrough", 0), 17: ("clear", 0)})
    golden_fd("fd_brir_2src_b128", 360, 300, 128, 10, 302, [(30, 0), (-45.5, 0)], is_hrir=False)
    golden_fd("fd_hrir_3src_b32", 360, 40, 32, 12, 303, [(0, 0), (90, 0), (200.25, 0)])

    # OLS: benchmark-like 2-ch, partitioned
    golden_ols("ols_2ch_1024_b256", syn.decaying_noise_irs((2, 1024), 201), 256, 8, 202)
    golden_ols("ols_2ch_dirac_b128", np.eye(2, 384, dtype=np.float32), 128, 5, 203)
    golden_ols("ols_hpcf_1024_b64", syn.decaying_noise_irs((2, 1024), 204), 64, 24, 205,
               is_hpcf=True, events={6: ("passthrough", 1), 9: ("passthrough", 0),
                                     20: ("clear", 0)})
    golden_ols("ols_arir_mono_6ch_b64", syn.decaying_noise_irs((6, 200), 206), 64, 10, 207,
               mono_in=True)
    golden_ols("ols_short_b256", syn.decaying_noise_irs((3, 100), 208), 256, 4, 209)

    # block delay line of the JACK clients (row f4)
    golden_delay("delay_3ch_b64")
    golden_levels("levels_4ch_b256")
    # compensation bookkeeping (row f1)
    golden_compensation()
    # the production chain ARIR -> renderer -> HPCF with the reference's own classes
    golden_production_chain()


if __name__ == "__main__":
    main()
