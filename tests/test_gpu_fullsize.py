"""GPU checks at the BASELINE.json sizes (1024 batched em32 streams; wide arrays with many streams):
spot checks of single streams against the numpy oracle plus size-independent properties (stream
independence, linearity).  The wide-array cases run the tensor-core analysis kernel over many
tiles per persistent CTA, which the single-stream golden fixtures cannot exercise."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _oracle_for(tb):
    from oracle import sh_chain as oc

    cfg = tb["sh_config"]
    return oc.ShOracle(cfg.sh_m, cfg.sh_bases_weighted, tb["hrir_set"].get_filter_blocks_nm(),
                       tb["B"], tb["order"])


@pytest.mark.parametrize("workload,S,T", [("c2", 1024, 4), ("c5", 200, 3), ("c4", 301, 3), ("b4096", 24, 3)])
def test_spot_streams_match_oracle_at_scale(workload, S, T):
    from retisar_b200 import workloads as wl

    tb = wl.sh_tables(workload, hrir_dirs=300)
    conv = wl.make_renderer(tb, S)
    rng = np.random.default_rng(S)
    C, B = tb["C"], tb["B"]
    probes = sorted({0, S - 1, int(rng.integers(1, S - 1)), int(rng.integers(1, S - 1))})
    oracles = {s: _oracle_for(tb) for s in probes}
    worst = 0.0
    for t in range(T):
        x = (rng.standard_normal((S, C, B)) * 0.1).astype(np.float32)
        yaw = rng.uniform(-720, 720, S).astype(np.float32)
        y = conv.filter_block(x, yaw)
        assert y.shape == (S, 2, B) and np.isfinite(y).all()
        for s, o in oracles.items():
            err = float(np.abs(y[s] - o.filter_block(x[s], yaw[s])).max())
            worst = max(worst, err)
            assert err <= TOL, (workload, t, s, err)
    print(workload, "S", S, "max abs err", worst)


def test_stream_independence_and_linearity_at_full_size():
    from retisar_b200 import workloads as wl

    tb = wl.sh_tables("c2", hrir_dirs=300)
    S, C, B = 1024, tb["C"], tb["B"]
    a, b = wl.make_renderer(tb, S), wl.make_renderer(tb, S)
    rng = np.random.default_rng(9)
    for t in range(3):
        x = (rng.standard_normal((S, C, B)) * 0.1).astype(np.float32)
        yaw = rng.uniform(0, 360, S).astype(np.float32)
        x[777], yaw[777] = x[5], yaw[5]          # two streams with identical input and yaw
        ya = a.filter_block(x, yaw)
        yb = b.filter_block(2.0 * x, yaw)        # scaling by 2 is exact in binary floating point
        assert np.array_equal(ya[777], ya[5])
        assert np.array_equal(yb, 2.0 * ya)


@pytest.mark.parametrize("workload,S", [("c2", 4100), ("c5", 2051), ("c4", 1030)])
def test_every_stream_of_a_large_batch_equals_the_first(workload, S):
    """All streams get the same input and yaw: every output must equal stream 0 bit for bit (each
    row of every tile of the persistent tensor-core kernel, through many mbarrier phases per CTA
    and a ragged last tile), and stream 0 must match the oracle."""
    from retisar_b200 import workloads as wl

    tb = wl.sh_tables(workload, hrir_dirs=200)
    conv = wl.make_renderer(tb, S)
    o = _oracle_for(tb)
    rng = np.random.default_rng(7)
    C, B = tb["C"], tb["B"]
    for t in range(3):
        x1 = (rng.standard_normal((C, B)) * 0.1).astype(np.float32)
        yaw1 = np.float32(rng.uniform(0, 360))
        y = conv.filter_block(np.broadcast_to(x1, (S, C, B)).copy(), np.full(S, yaw1, np.float32))
        assert np.array_equal(y, np.broadcast_to(y[0], y.shape)), (workload, t)
        assert np.abs(y[0] - o.filter_block(x1, yaw1)).max() <= TOL, (workload, t)
