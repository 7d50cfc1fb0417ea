"""Row a-X (extension): SH-domain rendering with a P-partition HRIR, CUDA vs the oracle defined by
composition of the reference's partition-queue semantics (oracle/sh_chain.py:
ShPartitionedOracle).  PARITY UNPINNED by the reference (it refuses P > 1); for P == 1 the oracle
is pinned to the reference (tests/test_oracle_vs_golden.py).  Bar: 1e-5 of full scale."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-5


def build(grid_name, order, B, taps, n_dirs, seed):
    from retisar_b200 import synthetic as syn
    from retisar_b200._filter_set import FilterSetMiro

    if grid_name == "rnd":
        az, co = syn.random_sphere_grid(40, seed)
        w = None
    else:
        az, co, w = syn.load_array_grid(grid_name)
    arr = FilterSetMiro.from_arrays(np.zeros((1, len(az), B), np.float32), az, co, w, order, False)
    arr.load(B, True, is_prevent_logging=True)
    haz, hco = syn.random_sphere_grid(n_dirs, seed + 1)
    hrir = FilterSetMiro.from_arrays(syn.decaying_noise_irs((n_dirs, 2, taps), seed + 2, tau=taps / 4),
                                     haz, hco, None, order, True)
    hrir.allow_partitions = True  # extension switch, see FilterSetMiro.calculate_filter_blocks_nm
    hrir.load(B, True, is_prevent_logging=True)
    return arr, hrir, len(az)


@pytest.mark.parametrize("tc", ["0", "1"])  # SIMT partition MAC / tensor-core partition MAC (forced)
@pytest.mark.parametrize("grid_name,order,B,taps,S", [
    ("em32", 4, 64, 200, 3),     # P = 4
    ("rnd", 3, 32, 90, 5),       # P = 3 (not a multiple of the register tile)
    ("le110", 8, 128, 700, 2),   # P = 6, order 8
    ("em32", 4, 256, 1100, 2),   # P = 5 at the headline block length
    ("em32", 4, 256, 1100, 131), # more than one 128-stream tile, ragged
])
def test_partitioned_sh_matches_composition_oracle(grid_name, order, B, taps, S, tc, monkeypatch):
    from oracle import sh_chain as oc
    from retisar_b200 import Convolver, _capi
    from retisar_b200 import synthetic as syn

    monkeypatch.setenv("RTB_PART_TC", tc)
    arr, hrir, C = build(grid_name, order, B, taps, 120, 31)
    conv = Convolver.create_instance_by_filter_set(hrir, B, [(0, 0)], None, n_streams=S)
    cfg = arr.get_sh_configuration()
    K, F = (order + 1) ** 2, B + 1
    conv.prepare_sh_processing(cfg, 0, None, comp_nm=np.ones((K, 1, F), np.complex64))
    hnm = hrir.get_filter_blocks_nm()
    P = hnm.shape[0]
    assert P == -(-taps // B) and conv._plan.P == P
    probes = list(range(S)) if S <= 8 else [0, 1, 63, 64, 127, 128, S - 1]
    oracles = {s: oc.ShPartitionedOracle(cfg.sh_m, cfg.sh_bases_weighted, hnm, B, order) for s in probes}
    T = 3 * P + 6
    x = syn.noise_blocks(T, S * C, B, 7).reshape(T, S, C, B)
    yaw = np.stack([syn.yaw_walk(T, 50 + s, step=25.0) for s in range(S)], axis=1)
    events = {P + 1: ("order", max(order - 2, 0)), P + 3: ("order", order), P + 4: ("crossfade", 0),
              2 * P + 2: ("crossfade", 1), 2 * P + 4: ("passthrough", 1), 2 * P + 5: ("passthrough", 0),
              3 * P + 2: ("clear", 0)}
    worst = 0.0
    for t in range(T):
        if t in events:
            kind, val = events[t]
            for o in oracles.values():
                if kind == "order":
                    o.set_order(val)
                elif kind == "crossfade":
                    o.set_crossfade(bool(val))
                elif kind == "passthrough":
                    o.is_passthrough = bool(val)
                else:
                    o.clear()
            if kind == "order":
                conv.update_sh_processing(val)
            elif kind == "crossfade":
                conv.set_crossfade(bool(val))
            elif kind == "passthrough":
                conv.set_passthrough(bool(val))
            else:
                conv._clear_buffers()
        y = conv.filter_block(x[t], yaw[t])
        ref = np.stack([o.filter_block(x[t, s], yaw[t, s]) for s, o in oracles.items()])
        err = np.abs(y[probes] - ref).max()
        worst = max(worst, err)
        assert err <= TOL, (t, err, events.get(t))
        if not conv._is_passthrough:
            want = _capi.RENDER_KERNEL_NAMES[6 if tc == "1" else 3]
            assert _capi.RENDER_KERNEL_NAMES[conv._plan.query(_capi.RTB_Q_RENDER_KERNEL)] == want
    print(grid_name, order, B, "P", P, "tc", tc, "max abs err", worst, "peak", np.abs(ref).max())


def test_partitioned_equals_long_convolution_property():
    """Size-independent property: with a fixed yaw and the cross-fade off, the partitioned
    renderer equals the single-partition renderer run on a longer block (same total filter)."""
    from oracle import sh_chain as oc
    from retisar_b200 import Convolver
    from retisar_b200 import synthetic as syn

    arr, hrir, C = build("em32", 4, 64, 256, 100, 77)   # P = 4 at B = 64
    conv = Convolver.create_instance_by_filter_set(hrir, 64, [(0, 0)], None)
    cfg = arr.get_sh_configuration()
    conv.prepare_sh_processing(cfg, 0, None, comp_nm=np.ones((25, 1, 65), np.complex64))
    conv.set_crossfade(False)
    arr2, hrir2, _ = build("em32", 4, 256, 256, 100, 77)  # same HRIR, P = 1 at B = 256
    hrir2.calculate_filter_blocks_fd(256)
    hrir2.calculate_filter_blocks_nm()
    o = oc.ShOracle(cfg.sh_m, cfg.sh_bases_weighted, hrir2.get_filter_blocks_nm(), 256, 4)
    o.set_crossfade(False)
    x = syn.noise_blocks(3, C, 256, 3)
    for t in range(3):
        y = np.concatenate([conv.filter_block(x[t][:, i * 64:(i + 1) * 64], [33.0]) for i in range(4)], axis=-1)
        ref = o.filter_block(x[t], 33.0)
        assert np.abs(y - ref).max() <= TOL


# ------------------------------------------------------------------------------------------------
# BASELINE config c3 at its real size: Lebedev-110, order 8, 4096-tap HRIR = 16 partitions, B = 256
# ------------------------------------------------------------------------------------------------
def _c3_tables(hrir_dirs=150):
    from retisar_b200 import workloads as wl

    return wl.sh_tables("c3", hrir_dirs=hrir_dirs)


def test_c3_full_size_events_match_composition_oracle():
    """le110 / N = 8 / P = 16 / B = 256 with S = 2051 streams (ragged last tile of every kernel of
    the partitioned path), 3P + 6 blocks with order / cross-fade / passthrough / clear events.
    Reference semantics restated by the oracle: queue roll `_convolver.py:645-665`, time-varying
    multiply `:838-866`, the TODO sketch `:1265-1269`.  Inputs live on the device; the probe
    streams are copied back for the per-stream numpy oracle."""
    import torch

    from oracle import sh_chain as oc
    from retisar_b200 import workloads as wl

    S = 2051
    tb = _c3_tables()
    conv = wl.make_renderer(tb, S)
    hnm = tb["hrir_set"].get_filter_blocks_nm()
    cfg = tb["sh_config"]
    B, C, order = tb["B"], tb["C"], tb["order"]
    P = hnm.shape[0]
    assert (P, B, C, order) == (16, 256, 110, 8) and conv._plan.P == 16
    probes = [0, 1, 127, 128, 1023, 2047, 2048, S - 1]
    oracles = {s: oc.ShPartitionedOracle(cfg.sh_m, cfg.sh_bases_weighted, hnm, B, order) for s in probes}
    T = 3 * P + 6
    events = {P + 1: ("order", 5), P + 3: ("order", order), P + 4: ("crossfade", 0),
              2 * P + 2: ("crossfade", 1), 2 * P + 4: ("passthrough", 1), 2 * P + 5: ("passthrough", 0),
              3 * P + 2: ("clear", 0)}
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(33)
    yaw = torch.rand((S,), generator=gen, device=dev) * 360.0
    worst, peak = 0.0, 0.0
    idx = torch.as_tensor(probes, device=dev)
    for t in range(T):
        if t in events:
            kind, val = events[t]
            for o in oracles.values():
                if kind == "order":
                    o.set_order(val)
                elif kind == "crossfade":
                    o.set_crossfade(bool(val))
                elif kind == "passthrough":
                    o.is_passthrough = bool(val)
                else:
                    o.clear()
            if kind == "order":
                conv.update_sh_processing(val)
            elif kind == "crossfade":
                conv.set_crossfade(bool(val))
            elif kind == "passthrough":
                conv.set_passthrough(bool(val))
            else:
                conv._clear_buffers()
        x = torch.randn((S, C, B), generator=gen, device=dev) * 0.1
        yaw = yaw + (torch.rand((S,), generator=gen, device=dev) - 0.5) * 40.0
        y = conv.filter_block(x, yaw)
        xs, ys, yaws = x[idx].cpu().numpy(), y[idx].cpu().numpy(), yaw[idx].cpu().numpy()
        assert bool(torch.isfinite(y).all())
        for i, s in enumerate(probes):
            ref = oracles[s].filter_block(xs[i], yaws[i])
            err = float(np.abs(ys[i] - ref).max())
            worst, peak = max(worst, err), max(peak, float(np.abs(ref).max()))
            assert err <= TOL, (t, s, err, events.get(t))
    print(f"c3 full size: S {S} P {P} max abs err {worst:.3g} (peak {peak:.3g}, rel {worst / peak:.3g})")
    assert worst <= 2e-6, worst  # regression bar, well inside the 1e-5 contract


def test_c3_p16_fixed_yaw_equals_reference_pinned_oracle_on_long_block():
    """Pin that does not depend on the composition oracle: with a fixed yaw and the cross-fade off
    the 16-partition renderer at B = 256 must equal the single-partition renderer at B = 4096 --
    and that one is `ShOracle`, pinned bit for bit to the unmodified reference."""
    from oracle import sh_chain as oc
    from retisar_b200 import synthetic as syn
    from retisar_b200 import workloads as wl
    from retisar_b200._filter_set import FilterSetMiro

    tb = _c3_tables()
    S = 130  # more than one 128-row tile
    conv = wl.make_renderer(tb, S)
    conv.set_crossfade(False)
    cfg, C, order = tb["sh_config"], tb["C"], tb["order"]
    long_b = 4096
    irs = tb["hrir_set"]._irs_td[:, :, :long_b]
    grid = tb["hrir_set"]._irs_grid
    hr2 = FilterSetMiro.from_arrays(irs, grid.azimuth, grid.colatitude, None, order, True)
    hr2.load(long_b, True, is_prevent_logging=True)
    hr2.calculate_filter_blocks_fd(long_b)
    hr2.calculate_filter_blocks_nm()
    assert hr2.get_filter_blocks_nm().shape[0] == 1
    yaws = np.linspace(-170.0, 185.0, S).astype(np.float32)
    probes = [0, 64, 128, S - 1]
    oracles = {}
    for s in probes:
        o = oc.ShOracle(cfg.sh_m, cfg.sh_bases_weighted, hr2.get_filter_blocks_nm(), long_b, order)
        o.set_crossfade(False)
        oracles[s] = o
    rng = np.random.default_rng(5)
    worst = 0.0
    for t in range(2):
        x = (rng.standard_normal((S, C, long_b)) * 0.1).astype(np.float32)
        y = np.concatenate([conv.filter_block(np.ascontiguousarray(x[:, :, i * 256:(i + 1) * 256]), yaws)
                            for i in range(long_b // 256)], axis=-1)
        for s, o in oracles.items():
            ref = o.filter_block(x[s], yaws[s])
            err = float(np.abs(y[s] - ref).max())
            worst = max(worst, err)
            assert err <= TOL, (t, s, err)
    print("c3 P = 16 vs reference-pinned B = 4096 oracle: max abs err", worst)


def test_partitioned_tensor_core_path_beyond_one_launch():
    """More than 128 stream tiles (16 384 streams): the tensor-core partition MAC runs as several
    launches over stream ranges; probes in the first and in the second launch, ragged last tile."""
    import torch

    from oracle import sh_chain as oc
    from retisar_b200 import Convolver, _capi

    S, B, order = 16384 + 200, 64, 4
    arr, hrir, C = build("em32", order, B, 2 * B, 100, 91)   # P = 2
    conv = Convolver.create_instance_by_filter_set(hrir, B, [(0, 0)], None, n_streams=S)
    cfg = arr.get_sh_configuration()
    conv.prepare_sh_processing(cfg, 0, None, comp_nm=np.ones(((order + 1) ** 2, 1, B + 1), np.complex64))
    hnm = hrir.get_filter_blocks_nm()
    probes = [0, 16383, 16384, 16384 + 127, S - 1]
    oracles = {s: oc.ShPartitionedOracle(cfg.sh_m, cfg.sh_bases_weighted, hnm, B, order) for s in probes}
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(8)
    idx = torch.as_tensor(probes, device=dev)
    for t in range(6):
        x = torch.randn((S, C, B), generator=gen, device=dev) * 0.1
        yaw = torch.rand((S,), generator=gen, device=dev) * 360.0
        y = conv.filter_block(x, yaw)
        assert _capi.RENDER_KERNEL_NAMES[conv._plan.query(_capi.RTB_Q_RENDER_KERNEL)] == "sh_partition_tc_kernel"
        xs, ys, yaws = x[idx].cpu().numpy(), y[idx].cpu().numpy(), yaw[idx].cpu().numpy()
        for i, s in enumerate(probes):
            assert np.abs(ys[i] - oracles[s].filter_block(xs[i], yaws[i])).max() <= TOL, (t, s)
