"""GPU tests of the reference-shaped class API: factory dispatch, tracker
contract, controls, error behaviour, compensation path, copies, torch tensors."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def tables():
    from retisar_b200 import workloads as wl

    return wl.sh_tables("c2", hrir_dirs=150)


def _oracle(tables, hnm):
    from oracle import sh_chain as oc

    cfg = tables["sh_config"]
    return oc.ShOracle(cfg.sh_m, cfg.sh_bases_weighted, hnm, tables["B"], tables["order"])


def test_renderer_with_designed_compensation_and_tracker(tables):
    """prepare_sh_processing with SHT+SHF+MRF, yaw read from the shared tracker array, output is
    a fresh writable float32 [2; B] array; order update keeps signal state."""
    from retisar_b200 import Convolver, HeadTracker, config
    from retisar_b200._convolver import AdjustableShConvolver
    from retisar_b200 import synthetic as syn

    tracker = HeadTracker.create_shared_array()
    conv = Convolver.create_instance_by_filter_set(tables["hrir_set"], tables["B"], [(0, 0)], tracker)
    assert type(conv) is AdjustableShConvolver
    with pytest.raises(RuntimeError):
        conv.filter_block(np.zeros((32, 256), np.float32))  # not prepared yet
    conv.prepare_sh_processing(tables["sh_config"], 18,
                               "SPHERICAL_HARMONICS_TAPERING+SPHERICAL_HEAD_FILTER")
    conv.init_fft_optimize()
    assert conv.get_input_channel_count() == 32 and conv.get_output_channel_count() == 2
    hnm = conv._filter.get_filter_blocks_nm().copy()
    assert not np.allclose(hnm, 0)
    o = _oracle(tables, hnm)
    x = syn.noise_blocks(10, 32, 256, 5)
    yaw = syn.yaw_walk(10, 6)
    for t in range(10):
        if t == 4:
            assert conv.update_sh_processing(2) == 2
            o.set_order(2)
        if t == 7:
            assert conv.update_sh_processing(None) == 4
            o.set_order(4)
        if t == 8:
            assert conv.update_sh_processing(9) is None  # out of bounds: ignored
        tracker[HeadTracker.DataIndex.AZIM] = float(yaw[t])
        y = conv.filter_block(x[t])
        assert y.shape == (2, 256) and y.dtype == np.float32 and y.flags.writeable
        y *= 0.5  # the caller scales in place (`_jack_client.py:795`)
        ref = o.filter_block(x[t], yaw[t])
        assert np.abs(2 * y - ref).max() <= TOL, t
    # pause protocol
    config.IS_RUNNING.clear()
    assert conv.filter_block(None) is None  # clears buffers
    config.IS_RUNNING.set()
    o.clear()
    tracker[HeadTracker.DataIndex.AZIM] = 10.0
    assert np.abs(conv.filter_block(x[0]) - o.filter_block(x[0], 10.0)).max() <= TOL


def test_controls_and_copy(tables):
    from retisar_b200 import workloads as wl
    from copy import copy
    from retisar_b200 import synthetic as syn

    conv = wl.make_renderer(tables, 1)
    o = _oracle(tables, conv._filter.get_filter_blocks_nm())
    x = syn.noise_blocks(8, 32, 256, 9)
    assert conv.set_crossfade(False) is False
    o.set_crossfade(False)
    assert conv.set_passthrough() is True  # toggle
    o.is_passthrough = True
    y = conv.filter_block(x[0], [30.0])
    assert np.abs(y - o.filter_block(x[0], 30.0)).max() <= TOL
    assert np.abs(y - x[0][:2]).max() <= TOL  # passthrough = first two microphones
    assert conv.set_passthrough(False) is False
    o.is_passthrough = False
    assert conv.set_crossfade(None) is True
    o.set_crossfade(True)
    for t in range(1, 4):
        assert np.abs(conv.filter_block(x[t], [40.0 * t]) - o.filter_block(x[t], 40.0 * t)).max() <= TOL
    clone = copy(conv)  # independent state, same tables
    assert clone is not conv and clone._plan is not conv._plan
    o2 = _oracle(tables, conv._filter.get_filter_blocks_nm())
    assert np.abs(clone.filter_block(x[5], [12.0]) - o2.filter_block(x[5], 12.0)).max() <= TOL
    assert np.abs(conv.filter_block(x[4], [160.0]) - o.filter_block(x[4], 160.0)).max() <= TOL


def test_errors_match_reference(tables):
    from retisar_b200 import Convolver
    from retisar_b200._filter_set import FilterSetMiro
    from retisar_b200.plans import ShPlan

    with pytest.raises(ValueError):  # no source positions (`_convolver.py:724-727`)
        Convolver.create_instance_by_filter_set(tables["hrir_set"], 256, None, None)
    with pytest.raises(NotImplementedError):  # more than one source (`:1036-1039`)
        Convolver.create_instance_by_filter_set(tables["hrir_set"], 256, [(0, 0), (90, 0)], None)
    with pytest.raises(NotImplementedError):  # partitions in the SH domain (`_filter_set.py:1136`)
        h = FilterSetMiro.from_arrays(np.zeros((10, 2, 600), np.float32), np.linspace(0, 6, 10),
                                      np.linspace(0.1, 3, 10), None, 1, True)
        h.load(256, True, is_prevent_logging=True)
        h.calculate_filter_blocks_fd(256)
        h.calculate_filter_blocks_nm()  # 3 blocks: refused unless `allow_partitions` is set
    with pytest.raises(NotImplementedError):
        ShPlan(1, 100, 32, 4)  # unsupported block length
    with pytest.raises(ValueError):
        ShPlan(1, 256, 32, -1)
    plan = ShPlan(1, 256, 32, 4)
    with pytest.raises(RuntimeError):  # tables missing (`_filter_set.py:1232-1233`)
        plan.process_host(np.zeros((1, 32, 256), np.float32), np.zeros(1, np.float32))
    with pytest.raises(ValueError):
        plan.set_order(7)


def test_misaligned_device_buffer_is_refused(tables):
    """The kernels use bulk copies and 16-byte loads: a device pointer off a 16-byte boundary is
    refused with the C ABI's invalid-argument status (ValueError), not faulted on."""
    import torch

    from retisar_b200 import workloads as wl

    S = 2
    conv = wl.make_renderer(tables, S)
    C, B = tables["C"], tables["B"]
    dev = torch.device("cuda", 0)
    flat = torch.zeros(S * C * B + 4, device=dev)
    yaw = torch.zeros(S, device=dev)
    out = torch.empty((S, 2, B), device=dev)
    with pytest.raises(ValueError):
        conv._plan.process_device(flat[1:1 + S * C * B].view(S, C, B), yaw, out, torch.cuda.current_stream())
    conv._plan.process_device(flat[4:4 + S * C * B].view(S, C, B), yaw, out, torch.cuda.current_stream())
    torch.cuda.synchronize()
    assert float(out.abs().max()) == 0.0


def test_torch_device_path_and_batched_ols(tables):
    import torch

    from oracle import sh_chain as oc
    from retisar_b200 import Convolver, FilterSet
    from retisar_b200 import workloads as wl

    S = 6
    conv = wl.make_renderer(tables, S)
    hnm = conv._filter.get_filter_blocks_nm()
    oracles = [_oracle(tables, hnm) for _ in range(S)]
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    for t in range(3):
        x = torch.randn((S, 32, 256), device="cuda", generator=g) * 0.1
        yaw = torch.rand((S,), device="cuda", generator=g) * 720 - 360
        y = conv.filter_block(x, yaw)
        assert y.is_cuda and y.shape == (S, 2, 256)
        ref = np.stack([o.filter_block(x[s].cpu().numpy(), float(yaw[s])) for s, o in enumerate(oracles)])
        assert np.abs(y.cpu().numpy() - ref).max() <= TOL
    # batched OLS convolver on the device, mono input broadcast
    irs = (np.random.default_rng(0).standard_normal((4, 300)) * 0.05).astype(np.float32)
    fs = FilterSet.create_instance_by_type(irs, "FIR_MULTICHANNEL")
    fs.load(block_length=64, is_single_precision=True, is_prevent_logging=True)
    ols = Convolver.create_instance_by_filter_set(fs, 64, n_streams=3)
    H = oc.filter_blocks_fd(oc.zero_pad(irs[None], 64), 64)
    refs = [oc.OlsOracle(H, 64) for _ in range(3)]
    for t in range(8):
        x = torch.randn((3, 1, 64), device="cuda", generator=g) * 0.1
        y = ols.filter_block(x)
        ref = np.stack([r.filter_block(x[s].cpu().numpy()) for s, r in enumerate(refs)])
        assert np.abs(y.cpu().numpy() - ref).max() <= TOL


@pytest.mark.parametrize("workload,dirs", [("c2", 2702), ("c3", 500), ("c4", 434)])
def test_sh_transform_of_the_hrir_set_on_the_device(workload, dirs):
    """`FilterSetMiro.calculate_filter_blocks_nm(device=0)` (tensor-core GEMM, `rtb_sh_transform_filter`)
    against the numpy product of the reference (`_filter_set.py:1093-1149`): 2702 directions (the
    reference's default HRIR set), 16 partitions (c3), order 12 = 338 analysis rows in two row
    chunks (c4).  Bar: 1e-5 of the largest coefficient (measured ~1e-6)."""
    from retisar_b200 import config
    from retisar_b200 import workloads as wl

    tb = wl.sh_tables(workload, hrir_dirs=dirs)
    h = tb["hrir_set"]
    h.calculate_filter_blocks_fd(tb["B"])
    h.calculate_filter_blocks_nm()
    ref = h.get_filter_blocks_nm().copy()
    h.calculate_filter_blocks_nm(device=0)
    got = h.get_filter_blocks_nm()
    assert got.shape == ref.shape and got.dtype == np.complex64
    err = float(np.abs(got - ref).max() / np.abs(ref).max())
    print(workload, got.shape, "relative error", err)
    assert err <= 1e-5
    # the renderer picks the device transform up through the config switch
    config.IS_SH_TRANSFORM_ON_DEVICE = True
    try:
        conv = wl.make_renderer(tb, 2)
        assert np.abs(tb["hrir_set"].get_filter_blocks_nm() - ref).max() / np.abs(ref).max() <= 1e-5
        x = (np.random.default_rng(0).standard_normal((2, tb["C"], tb["B"])) * 0.1).astype(np.float32)
        assert np.isfinite(conv.filter_block(x, [10.0, 200.0])).all()
    finally:
        config.IS_SH_TRANSFORM_ON_DEVICE = False


def test_async_host_blocks_and_block_runs_match_the_synchronous_path():
    """`submit_block` / `result()` (three staging slots, several blocks in flight, results collected
    out of submission order) and `rtb_sh_process_blocks` (a run of blocks enqueued by one call) give
    bit for bit what the synchronous `filter_block` gives, and that matches the oracle."""
    import torch

    from oracle import sh_chain as oc
    from retisar_b200 import workloads as wl

    tb = wl.sh_tables("c2", hrir_dirs=150)
    S, T, C, B = 37, 9, tb["C"], tb["B"]
    sync, asyn, runs = (wl.make_renderer(tb, S) for _ in range(3))
    rng = np.random.default_rng(21)
    x = (rng.standard_normal((T, S, C, B)) * 0.1).astype(np.float32)
    yaw = rng.uniform(0, 360, (T, S)).astype(np.float32)
    want = np.stack([sync.filter_block(x[t], yaw[t]).copy() for t in range(T)])
    cfg = tb["sh_config"]
    o = oc.ShOracle(cfg.sh_m, cfg.sh_bases_weighted, tb["hrir_set"].get_filter_blocks_nm(), B, tb["order"])
    for t in range(T):
        assert np.abs(want[t, 5] - o.filter_block(x[t, 5], yaw[t, 5])).max() <= 1e-5
    # five blocks in flight at once (more than the three staging slots), collected newest first
    tickets = [asyn.submit_block(x[t], yaw[t]) for t in range(5)]
    got = {t: tickets[t].result() for t in (4, 2, 0, 1, 3)}
    for t in range(5, T):  # feeder shape: submit t + 1, then collect t
        tickets.append(asyn.submit_block(x[t], yaw[t]))
        if t > 5:
            got[t - 1] = tickets[t - 1].result()
    got[T - 1] = tickets[T - 1].result()
    for t in range(T):
        assert np.array_equal(got[t], want[t]), t
    assert len({id(v) for v in got.values()}) == T  # every result is its own array
    # a run of blocks from a device-resident ring in one library call
    dev = torch.device("cuda", 0)
    ring = 4
    xr = torch.from_numpy(x[:ring]).to(dev).contiguous()
    yr = torch.from_numpy(yaw[:ring]).to(dev).contiguous()
    out = torch.empty((S, 2, B), device=dev)
    runs._plan.process_blocks_device(xr, yr, out, 0, 3, torch.cuda.current_stream())
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), want[2])
    runs._plan.process_blocks_device(xr, yr, out, 3, 1, torch.cuda.current_stream())
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), want[3])


def test_runs_of_blocks_replayed_as_cuda_graph_equal_single_blocks(monkeypatch):
    """`rtb_sh_process_blocks` replays a run that comes back with the same buffers, ring phase, length
    and plan state as one CUDA graph launch (recorded on the second occurrence).  Five identical runs
    and a sixth after a cross-fade toggle must equal block-by-block processing of a second plan; the
    launch counter keeps counting kernels."""
    import torch

    from retisar_b200 import _capi
    from retisar_b200 import workloads as wl

    tb = wl.sh_tables("c2", hrir_dirs=200)
    S, C, B, ring, count = 6, tb["C"], tb["B"], 4, 8
    rng = np.random.default_rng(33)
    x = (rng.standard_normal((ring, S, C, B)) * 0.1).astype(np.float32)
    yaw = rng.uniform(0, 360, (ring, S)).astype(np.float32)
    dev = torch.device("cuda", 0)
    xr, yr = torch.from_numpy(x).to(dev), torch.from_numpy(yaw).to(dev)
    st = torch.cuda.Stream(device=dev)
    results = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("RTB_GRAPH", mode)
        conv = wl.make_renderer(tb, S)
        out = torch.empty((S, 2, B), device=dev)
        outs, launches = [], []
        torch.cuda.synchronize()
        first = 0
        for run in range(13):
            if run == 5:
                conv.set_crossfade(False)
            if run == 6:
                conv.set_crossfade(True)
            if run == 8:    # lower order: the grouped tables are rebuilt (recorded runs must be dropped)
                conv._plan.set_order(2)
            if run == 10:   # back to the order the first recordings were made for
                conv._plan.set_order(tb["order"])
            l0 = conv._plan.query(_capi.RTB_Q_LAUNCHES)
            conv._plan.process_blocks_device(xr, yr, out, first, count, st)
            st.synchronize()
            outs.append(out.cpu().numpy().copy())
            launches.append(conv._plan.query(_capi.RTB_Q_LAUNCHES) - l0)
            first += count
        results[mode] = (outs, launches, conv._plan.query(_capi.RTB_Q_GRAPH_REPLAYS))
    for a, b in zip(results["1"][0], results["0"][0]):
        assert np.array_equal(a, b)
    assert results["1"][1] == results["0"][1] and results["1"][1][0] == 2 * count
    # runs 0-4 share a key (replayed from the third on), the later ones follow a state change each
    assert results["0"][2] == 0 and results["1"][2] >= 3
