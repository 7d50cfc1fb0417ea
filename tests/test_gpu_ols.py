"""GPU parity tests of the partitioned overlap-save convolver (row a8) through the C ABI and
through the reference-shaped OverlapSaveConvolver class.  Bar: 1e-5 of full scale."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu
TOL = 1e-5

OLS_CASES = ["ols_2ch_1024_b256", "ols_2ch_dirac_b128", "ols_hpcf_1024_b64",
             "ols_arir_mono_6ch_b64", "ols_short_b256"]


@pytest.mark.parametrize("name", OLS_CASES)
def test_ols_plan_matches_reference_golden(name):
    from retisar_b200.plans import OlsPlan

    g = load_golden(name)
    B = int(g["B"])
    H = g["Hfd"][:, 0]  # [P, Cout, F]
    n_in = g["x"].shape[1]
    plan = OlsPlan(1, B, n_in, H.shape[1], H.shape[0])
    plan.set_filter(H)
    worst = 0.0
    for t in range(g["x"].shape[0]):
        kind, val = g["events"][t]
        if kind == 3:
            plan.set_passthrough(bool(val))
        elif kind == 4:
            plan.reset()
        y = plan.process_host(g["x"][t][None])[0]
        err = np.abs(y - g["y"][t]).max()
        worst = max(worst, err)
        assert err <= TOL, (name, t, err)
    print(name, "max abs err", worst)


@pytest.mark.parametrize("name", ["ols_2ch_1024_b256", "ols_arir_mono_6ch_b64"])
def test_overlap_save_convolver_class(name):
    """Same through FilterSet.create_instance_by_type + Convolver.create_instance_by_filter_set,
    the way `_run_benchmark` builds its convolvers (`_jack_renderer.py:478-499`)."""
    from retisar_b200 import Convolver, FilterSet
    from retisar_b200._convolver import OverlapSaveConvolver

    g = load_golden(name)
    B = int(g["B"])
    fs = FilterSet.create_instance_by_type(g["irs"], "FIR_MULTICHANNEL")
    fs.load(block_length=B, is_single_precision=True, is_prevent_logging=True)
    conv = Convolver.create_instance_by_filter_set(fs, B)
    assert type(conv) is OverlapSaveConvolver
    assert conv.get_output_channel_count() == g["irs"].shape[0]
    for t in range(g["x"].shape[0]):
        y = conv.filter_block(g["x"][t])
        assert y.shape == g["y"][t].shape and y.flags.writeable
        assert np.abs(y - g["y"][t]).max() <= TOL
    assert conv.filter_block(None) is None


def test_ols_batched_streams_long_filter():
    """S streams x 3 channels, 20 partitions: equals direct convolution (size-independent check)."""
    from retisar_b200.plans import OlsPlan

    rng = np.random.default_rng(3)
    S, C, B, P, T = 5, 3, 64, 20, 30
    irs = (rng.standard_normal((C, P * B)) * 0.05 * np.exp(-np.arange(P * B) / 300.0)).astype(np.float32)
    H = np.stack([np.fft.rfft(irs[:, p * B:(p + 1) * B], 2 * B) for p in range(P)]).astype(np.complex64)
    x = (rng.standard_normal((T, S, C, B)) * 0.1).astype(np.float32)
    plan = OlsPlan(S, B, C, C, P)
    plan.set_filter(H)
    y = np.concatenate([plan.process_host(x[t]) for t in range(T)], axis=-1)
    xs = np.concatenate(list(x), axis=-1).astype(np.float64)
    for s in range(S):
        for c in range(C):
            ref = np.convolve(xs[s, c], irs[c].astype(np.float64))[: T * B]
            assert np.abs(y[s, c] - ref).max() <= TOL
