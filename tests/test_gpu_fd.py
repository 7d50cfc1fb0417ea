"""GPU parity tests of the HRIR / BRIR switching renderer (row f3, AdjustableFdConvolver) against
fixtures produced by the unmodified reference.  Bar: 1e-5 of full scale."""
import numpy as np
import pytest

from conftest import check_regression_bar, load_golden

pytestmark = pytest.mark.gpu
TOL = 1e-5
FD_CASES = ["fd_hrir_1src_b64", "fd_brir_2src_b128", "fd_hrir_3src_b32"]


def filter_set(g):
    from retisar_b200 import synthetic as syn
    from retisar_b200._filter_set import FilterSetSsr

    irs = syn.decaying_noise_irs(tuple(g["irs_shape"]), int(g["irs_seed"]))
    assert abs(np.abs(irs.astype(np.float64)).sum() - float(g["irs_sum"])) < 1e-6
    fs = FilterSetSsr.from_arrays(irs, is_hrir=bool(g["is_hrir"]))
    fs.load(block_length=int(g["B"]), is_single_precision=True, is_prevent_logging=True)
    return fs


# the launch variants of the renderer: head kernel + queue kernel (default), and the fused kernel in
# its two shapes (RTB_FD_SPLIT / RTB_FD_DENSE are read when the plan is created)
FD_LAUNCHES = {"split": {"RTB_FD_SPLIT": "1"}, "fused": {"RTB_FD_SPLIT": "0", "RTB_FD_DENSE": "0"},
               "fused_dense": {"RTB_FD_SPLIT": "0", "RTB_FD_DENSE": "1"}}


@pytest.mark.parametrize("launch", list(FD_LAUNCHES))
@pytest.mark.parametrize("name", FD_CASES)
def test_fd_convolver_matches_reference_golden(name, launch, monkeypatch):
    from retisar_b200 import Convolver, HeadTracker
    from retisar_b200._convolver import AdjustableFdConvolver

    for k, v in FD_LAUNCHES[launch].items():
        monkeypatch.setenv(k, v)
    g = load_golden(name)
    B = int(g["B"])
    tracker = HeadTracker.create_shared_array()
    conv = Convolver.create_instance_by_filter_set(
        filter_set(g), B, [tuple(map(float, s)) for s in g["sources"]], tracker)
    assert type(conv) is AdjustableFdConvolver
    assert conv.get_input_channel_count() == len(g["sources"]) and conv.get_output_channel_count() == 2
    worst = 0.0
    for t in range(g["x"].shape[0]):
        kind, val = g["events"][t]
        if kind == 2:
            conv.set_crossfade(bool(val))
        elif kind == 3:
            conv.set_passthrough(bool(val))
        elif kind == 4:
            conv._clear_buffers()
        tracker[HeadTracker.DataIndex.AZIM] = float(g["yaw"][t])
        y = conv.filter_block(g["x"][t])
        err = np.abs(y - g["y"][t]).max()
        worst = max(worst, err)
        assert err <= TOL, (name, t, err)
    check_regression_bar(float(worst), float(np.abs(g["y"]).max()), name)


@pytest.mark.parametrize("launch", list(FD_LAUNCHES))
@pytest.mark.parametrize("n_src", [2, 5])  # 5 sources: three source pairs, the queue kernel's generic path
def test_fd_batched_streams_vs_oracle(n_src, launch, monkeypatch):
    from oracle import sh_chain as oc
    from retisar_b200 import Convolver

    for k, v in FD_LAUNCHES[launch].items():
        monkeypatch.setenv(k, v)
    g = load_golden("fd_brir_2src_b128")
    B, S, T = int(g["B"]), 5, 14
    fs = filter_set(g)
    sources = [tuple(map(float, s)) for s in g["sources"]]
    if n_src > 2:
        sources = sources + [(75.0, 0.0), (-120.0, 0.0), (200.0, 0.0)][: n_src - 2]
    conv = Convolver.create_instance_by_filter_set(fs, B, sources, None, n_streams=S)
    oracles = [oc.FdOracle(fs._irs_blocks_fd, B, np.array(sources), bool(g["is_hrir"])) for _ in range(S)]
    rng = np.random.default_rng(4)
    # state changes between blocks: the split launch leaves consumed queue heads for the next block
    # to overwrite, so every change of launch variant (passthrough) or of the queues in use
    # (cross-fade) in mid-stream has to keep the queues consistent
    events = {4: ("passthrough", True), 5: ("passthrough", False), 7: ("crossfade", False),
              9: ("crossfade", True), 11: ("passthrough", True), 12: ("passthrough", False)}
    for t in range(T):
        if t in events:
            what, state = events[t]
            getattr(conv, "set_" + what)(state)
            for o in oracles:
                if what == "crossfade":
                    o.set_crossfade(state)
                else:
                    o.is_passthrough = state
        x = (rng.standard_normal((S, n_src, B)) * 0.1).astype(np.float32)
        yaw = rng.uniform(-500, 500, S).astype(np.float32)
        y = conv.filter_block(x, yaw)
        ref = np.stack([o.filter_block(x[s], yaw[s]) for s, o in enumerate(oracles)])
        assert np.abs(y - ref).max() <= TOL, t


def test_fd_plan_refuses_fewer_than_360_directions():
    """The direction index is int(-azimuth) on a whole-degree grid: fewer than 360 directions would
    index out of bounds on the device (the reference raises IndexError at run time)."""
    from retisar_b200.plans import FdPlan

    with pytest.raises(ValueError, match="360 directions"):
        FdPlan(2, 64, 1, 72, 2)
