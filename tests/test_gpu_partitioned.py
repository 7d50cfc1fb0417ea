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


@pytest.mark.parametrize("grid_name,order,B,taps,S", [
    ("em32", 4, 64, 200, 3),     # P = 4
    ("rnd", 3, 32, 90, 5),       # P = 3 (not a multiple of the register tile)
    ("le110", 8, 128, 700, 2),   # P = 6, order 8
    ("em32", 4, 256, 1100, 2),   # P = 5 at the headline block length
])
def test_partitioned_sh_matches_composition_oracle(grid_name, order, B, taps, S):
    from oracle import sh_chain as oc
    from retisar_b200 import Convolver
    from retisar_b200 import synthetic as syn

    arr, hrir, C = build(grid_name, order, B, taps, 120, 31)
    conv = Convolver.create_instance_by_filter_set(hrir, B, [(0, 0)], None, n_streams=S)
    cfg = arr.get_sh_configuration()
    K, F = (order + 1) ** 2, B + 1
    conv.prepare_sh_processing(cfg, 0, None, comp_nm=np.ones((K, 1, F), np.complex64))
    hnm = hrir.get_filter_blocks_nm()
    P = hnm.shape[0]
    assert P == -(-taps // B) and conv._plan.P == P
    oracles = [oc.ShPartitionedOracle(cfg.sh_m, cfg.sh_bases_weighted, hnm, B, order) for _ in range(S)]
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
            for o in oracles:
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
        ref = np.stack([o.filter_block(x[t, s], yaw[t, s]) for s, o in enumerate(oracles)])
        err = np.abs(y - ref).max()
        worst = max(worst, err)
        assert err <= TOL, (t, err, events.get(t))
    print(grid_name, order, B, "P", P, "max abs err", worst, "peak", np.abs(ref).max())


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
