"""GPU parity tests at the large block lengths (512 ... 4096; 4096 is the reference's default,
`ReTiSAR/config.py:9`, and `_run_benchmark.py:38-46` sweeps 128 ... 4096).  These sizes run the
CTA-wide FFT (one transform per thread block) instead of the warp-level one, through the same
kernels.  Checked against the numpy oracle (pinned bit-exact to the reference by
tests/test_oracle_vs_golden.py) on seeded inputs, and against direct convolution.
Bar: 1e-5 of full scale."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-5
BLOCKS = [512, 1024, 2048, 4096]


def _sh_tables(B, order, n_ch, taps, seed, symmetric=True):
    from oracle import sh_chain as oc
    from retisar_b200 import synthetic as syn

    az, co = syn.random_sphere_grid(n_ch, seed)
    sh_m, Yw = oc.sh_configuration(order, az, co, None)[:2]
    if not symmetric:  # arbitrary analysis matrix: the general (2K-row) mode
        rng = np.random.default_rng(seed + 50)
        Yw = (Yw + 0.05 * (rng.standard_normal(Yw.shape) + 1j * rng.standard_normal(Yw.shape))).astype(np.complex64)
    haz, hco = syn.random_sphere_grid(120, seed + 1)
    hrir = syn.decaying_noise_irs((120, 2, taps), seed + 2)
    h_fd = oc.filter_blocks_fd(oc.zero_pad(hrir, B), B)
    Yh = oc.sh_configuration(order, haz, hco, None)[1]
    Hnm = oc.filter_blocks_nm(h_fd, Yh)
    return sh_m, Yw, Hnm


@pytest.mark.parametrize("B", BLOCKS)
def test_sh_chain_large_block(B):
    from oracle import sh_chain as oc
    from retisar_b200.plans import ShPlan

    order, C, S, T = 3, 20, 3, 5
    sh_m, Yw, Hnm = _sh_tables(B, order, C, 200, 11 + B)
    rev = oc.reverse_mn_ids(order)
    w_in, w_out = oc.crossfade_windows(B)
    plan = ShPlan(S, B, C, order)
    plan.set_tables(Yw, Hnm, sh_m, rev, w_in, w_out)
    oracles = [oc.ShOracle(sh_m, Yw, Hnm, B, order) for _ in range(S)]
    rng = np.random.default_rng(B)
    worst = 0.0
    for t in range(T):
        if t == 3:  # order lowering: one block with different orders for the two yaws
            plan.set_order(2)
            for o in oracles:
                o.set_order(2)
        x = (rng.standard_normal((S, C, B)) * 0.1).astype(np.float32)
        yaw = rng.uniform(-400, 400, S).astype(np.float32)
        y = plan.process_host(x, yaw)
        ref = np.stack([o.filter_block(x[s], yaw[s]) for s, o in enumerate(oracles)])
        worst = max(worst, float(np.abs(y - ref).max()))
        assert np.abs(y - ref).max() <= TOL, (B, t, np.abs(y - ref).max())
    print("B", B, "max abs err", worst, "peak", float(np.abs(ref).max()))


def test_sh_chain_large_block_general_mode_and_controls():
    from oracle import sh_chain as oc
    from retisar_b200.plans import ShPlan

    B, order, C, S = 1024, 2, 12, 2
    sh_m, Yw, Hnm = _sh_tables(B, order, C, 300, 5, symmetric=False)
    plan = ShPlan(S, B, C, order)
    plan.set_tables(Yw, Hnm, sh_m, oc.reverse_mn_ids(order), *oc.crossfade_windows(B))
    oracles = [oc.ShOracle(sh_m, Yw, Hnm, B, order) for _ in range(S)]
    rng = np.random.default_rng(2)
    for t in range(7):
        if t == 2:
            plan.set_crossfade(False)
            for o in oracles:
                o.set_crossfade(False)
        if t == 4:
            plan.set_crossfade(True)
            for o in oracles:
                o.set_crossfade(True)
        if t in (5, 6):
            plan.set_passthrough(t == 5)
            for o in oracles:
                o.is_passthrough = t == 5
        x = (rng.standard_normal((S, C, B)) * 0.1).astype(np.float32)
        yaw = rng.uniform(0, 360, S).astype(np.float32)
        y = plan.process_host(x, yaw)
        ref = np.stack([o.filter_block(x[s], yaw[s]) for s, o in enumerate(oracles)])
        assert np.abs(y - ref).max() <= TOL, (t, np.abs(y - ref).max())


@pytest.mark.parametrize("B", BLOCKS)
def test_ols_large_block(B):
    """2-in/2-out and mono-in/3-out partitioned convolvers: oracle and direct convolution."""
    from oracle import sh_chain as oc
    from retisar_b200.plans import OlsPlan

    rng = np.random.default_rng(B + 1)
    for (cin, cout, P) in ((2, 2, 3), (1, 3, 2)):
        S, T = 2, 5
        irs = (rng.standard_normal((cout, P * B)) * 0.05 * np.exp(-np.arange(P * B) / 900.0)).astype(np.float32)
        H = np.stack([np.fft.rfft(irs[:, p * B:(p + 1) * B], 2 * B) for p in range(P)]).astype(np.complex64)
        plan = OlsPlan(S, B, cin, cout, P)
        plan.set_filter(H)
        oracles = [oc.OlsOracle(H[:, None], B) for _ in range(S)]
        x = (rng.standard_normal((T, S, cin, B)) * 0.1).astype(np.float32)
        ys = []
        for t in range(T):
            y = plan.process_host(x[t])
            ref = np.stack([o.filter_block(x[t, s]) for s, o in enumerate(oracles)])
            assert np.abs(y - ref).max() <= TOL, (B, cin, t, np.abs(y - ref).max())
            ys.append(y)
        y = np.concatenate(ys, axis=-1)
        xs = np.concatenate(list(x), axis=-1).astype(np.float64)
        for s in range(S):
            for c in range(cout):
                ref = np.convolve(xs[s, c if cin > 1 else 0], irs[c].astype(np.float64))[: T * B]
                assert np.abs(y[s, c] - ref).max() <= TOL


@pytest.mark.parametrize("B", [512, 4096])
def test_fd_large_block(B):
    from oracle import sh_chain as oc
    from retisar_b200 import Convolver
    from retisar_b200 import synthetic as syn
    from retisar_b200._filter_set import FilterSetSsr

    irs = syn.decaying_noise_irs((360, 2, B + 100), 9)
    fs = FilterSetSsr.from_arrays(irs, is_hrir=True)
    fs.load(block_length=B, is_single_precision=True, is_prevent_logging=True)
    sources = [(30.0, 0.0), (-45.0, 0.0), (110.0, 0.0)]
    S = 2
    conv = Convolver.create_instance_by_filter_set(fs, B, sources, None, n_streams=S)
    oracles = [oc.FdOracle(fs._irs_blocks_fd, B, np.array(sources), True) for _ in range(S)]
    rng = np.random.default_rng(B)
    for t in range(5):
        x = (rng.standard_normal((S, 3, B)) * 0.1).astype(np.float32)
        yaw = rng.uniform(-500, 500, S).astype(np.float32)
        y = conv.filter_block(x, yaw)
        ref = np.stack([o.filter_block(x[s], yaw[s]) for s, o in enumerate(oracles)])
        assert np.abs(y - ref).max() <= TOL, (B, t, np.abs(y - ref).max())


def test_reference_default_block_length_through_the_classes():
    """The reference's default configuration (block 4096) through FilterSet / Convolver."""
    from retisar_b200 import Convolver, FilterSet
    from retisar_b200._convolver import OverlapSaveConvolver

    B = 4096
    rng = np.random.default_rng(0)
    irs = (rng.standard_normal((2, 3000)) * 0.05).astype(np.float32)
    fs = FilterSet.create_instance_by_type(irs, "FIR_MULTICHANNEL")
    fs.load(block_length=B, is_single_precision=True, is_prevent_logging=True)
    conv = Convolver.create_instance_by_filter_set(fs, B)
    assert type(conv) is OverlapSaveConvolver
    x = (rng.standard_normal((3, 2, B)) * 0.1).astype(np.float32)
    y = np.concatenate([conv.filter_block(x[t]) for t in range(3)], axis=-1)
    xs = np.concatenate(list(x), axis=-1).astype(np.float64)
    for c in range(2):
        ref = np.convolve(xs[c], irs[c].astype(np.float64))[: 3 * B]
        assert np.abs(y[c] - ref).max() <= TOL


@pytest.mark.parametrize("order,C,B", [(5, 30, 128), (6, 32, 256), (7, 24, 64)])
def test_narrow_array_more_rows_than_one_mma_pass(order, C, B):
    """Narrow arrays (C <= 32) with more than 32 real analysis rows: the warp-level tensor-core
    analysis kernel streams its tiles once per chunk of 32 rows."""
    from oracle import sh_chain as oc
    from retisar_b200 import _capi
    from retisar_b200.plans import ShPlan

    S, T = 5, 4
    sh_m, Yw, Hnm = _sh_tables(B, order, C, min(100, B), 300 + order)
    plan = ShPlan(S, B, C, order)
    plan.set_tables(Yw, Hnm, sh_m, oc.reverse_mn_ids(order), *oc.crossfade_windows(B))
    assert plan.query(_capi.RTB_Q_ROWS) > 32
    oracles = [oc.ShOracle(sh_m, Yw, Hnm, B, order) for _ in range(S)]
    rng = np.random.default_rng(order)
    for t in range(T):
        x = (rng.standard_normal((S, C, B)) * 0.1).astype(np.float32)
        yaw = rng.uniform(0, 360, S).astype(np.float32)
        y = plan.process_host(x, yaw)
        ref = np.stack([o.filter_block(x[s], yaw[s]) for s, o in enumerate(oracles)])
        assert np.abs(y - ref).max() <= TOL, (order, t, np.abs(y - ref).max())
