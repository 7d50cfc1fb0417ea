"""GPU parity tests of the partitioned overlap-save convolver (row a8) through the C ABI and
through the reference-shaped OverlapSaveConvolver class.  Bar: 1e-5 of full scale."""
import numpy as np
import pytest

from conftest import check_regression_bar, load_golden

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
    check_regression_bar(float(worst), float(np.abs(g["y"]).max()), name)


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


@pytest.mark.parametrize("B,Cout,P,S", [(256, 33, 3, 5), (32, 40, 5, 130), (64, 120, 20, 3), (128, 32, 1, 2)])
def test_mono_in_many_out_tensor_core_path(B, Cout, P, S):
    """Mono input broadcast into >= 32 filter channels (the ARIR pre-renderer shape) runs as a
    per-bin GEMM on the tensor cores: odd channel counts, partition counts that are not a multiple
    of the K chunk, stream counts on both sides of one M = 128 tile, every block length of the path,
    with passthrough blocks and a clear in between -- block by block against the pinned oracle."""
    from oracle import sh_chain as oc
    from retisar_b200.plans import OlsPlan

    rng = np.random.default_rng(B + Cout)
    T = 3 * P + 6
    irs = (rng.standard_normal((Cout, P * B)) * 0.05 * np.exp(-np.arange(P * B) / (0.3 * P * B))).astype(np.float32)
    H = oc.filter_blocks_fd(irs[None], B)  # [P, 1, Cout, F]
    plan = OlsPlan(S, B, 1, Cout, P)
    plan.set_filter(H[:, 0])
    oracles = [oc.OlsOracle(H.copy(), B) for _ in range(S)]
    x = (rng.standard_normal((T, S, 1, B)) * 0.3).astype(np.float32)
    worst = 0.0
    for t in range(T):
        passthrough = t in (P + 1, P + 2)
        plan.set_passthrough(passthrough)
        if t == 2 * P + 3:
            plan.reset()
        y = plan.process_host(x[t])
        for s, o in enumerate(oracles):
            o.is_passthrough = passthrough
            if t == 2 * P + 3:
                o.clear()
            worst = max(worst, float(np.abs(y[s] - o.filter_block(x[t, s])).max()))
    assert worst <= TOL, worst


def test_reset_is_ordered_against_blocks_in_flight_and_mono_after_full_width():
    """`filter_block(None)` / `_clear_buffers` right behind a device block on a non-blocking stream
    must clear the state that block wrote (the reset drains the device first), and a mono block
    after full-width ones is broadcast into the same state like `_convolver.py:588-591`."""
    import torch

    from oracle import sh_chain as oc
    from retisar_b200 import Convolver, FilterSet
    from retisar_b200 import synthetic as syn

    B, taps, n_ch, S = 128, 1000, 3, 64
    irs = syn.decaying_noise_irs((n_ch, taps), 3)
    fs = FilterSet.create_instance_by_type(irs, "FIR_MULTICHANNEL")
    fs.load(block_length=B, is_single_precision=True, is_prevent_logging=True)
    conv = Convolver.create_instance_by_filter_set(fs, B, n_streams=S)
    side = torch.cuda.Stream()
    rng = np.random.default_rng(2)
    x = (rng.standard_normal((6, S, n_ch, B)) * 0.1).astype(np.float32)
    with torch.cuda.stream(side):
        for t in range(3):
            conv.filter_block(torch.from_numpy(x[t]).cuda())
        conv._clear_buffers()  # must wait for the three blocks above, then clear
        y = conv.filter_block(torch.from_numpy(x[3]).cuda())
    side.synchronize()
    o = oc.OlsOracle(fs.get_filter_blocks_fd(), B)
    ref = o.filter_block(x[3, 7])
    assert np.abs(y[7].cpu().numpy() - ref).max() <= 1e-5
    # mono block after a full-width one: broadcast, state kept
    xm = x[4][:, :1]
    ym = conv.filter_block(xm)
    refm = o.filter_block(xm[7])
    assert ym.shape == (S, n_ch, B) and np.abs(ym[7] - refm).max() <= 1e-5
    with pytest.raises(ValueError, match="input channel"):
        conv.filter_block(x[5][:, :2])
