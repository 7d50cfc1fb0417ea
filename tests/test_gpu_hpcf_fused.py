"""Row a-Y: headphone-compensation filter folded into the renderer's kernels
(`AdjustableShConvolver.set_hpcf`, csrc/hpcf_stage.cuh) against the reference's production
topology renderer -> separate HPCF client (`_run.py:290-338`), i.e. `OlsOracle` applied to the
output of `ShOracle` (`_convolver.py:524-571` after `:1231-1316`).  Bar: 1e-5 of full scale."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _hpcf_set(B, taps, seed):
    from retisar_b200 import FilterSet
    from retisar_b200 import synthetic as syn

    irs = syn.decaying_noise_irs((2, taps), seed, scale=0.3)
    fs = FilterSet.create_instance_by_type(irs, "FIR_MULTICHANNEL")
    fs.load(block_length=B, is_single_precision=True, is_prevent_logging=True)
    fs.calculate_filter_blocks_fd(B)
    return fs


def _run(conv, sh_oracles, hp_oracles, x, yaw, events):
    worst = 0.0
    for t in range(x.shape[0]):
        if t in events:
            kind, val = events[t]
            for o in sh_oracles:
                if kind == "order":
                    o.set_order(val)
                elif kind == "crossfade":
                    o.set_crossfade(bool(val))
                elif kind == "passthrough":
                    o.is_passthrough = bool(val)
                else:
                    o.clear()
            if kind == "clear":  # both clients clear their buffers when the pipeline is paused
                for o in hp_oracles:
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
        ref = np.stack([hp.filter_block(o.filter_block(x[t, s], yaw[t, s]))
                        for s, (o, hp) in enumerate(zip(sh_oracles, hp_oracles))])
        err = float(np.abs(y - ref).max())
        worst = max(worst, err)
        assert err <= TOL, (t, err, events.get(t))
    return worst


@pytest.mark.parametrize("workload,S,hp_taps,partitioned", [
    ("c2", 6, 1024, False),    # m-grouped render kernel + fused stage (N2 = 512), P_hpcf = 4
    ("c2", 5, 100, False),     # single HPCF partition
    ("c4", 7, 1024, False),    # 64-sample blocks: two transforms per warp, term-wise kernel, P_hpcf = 16
    ("b4096", 2, 5000, False), # CTA-wide transforms
    ("c2", 3, 700, True),      # partitioned HRIR path: stand-alone stage behind the tail kernel
])
def test_fused_hpcf_equals_renderer_then_separate_convolver(workload, S, hp_taps, partitioned):
    from oracle import sh_chain as oc
    from retisar_b200 import synthetic as syn
    from retisar_b200 import workloads as wl
    from retisar_b200._filter_set import FilterSetMiro

    tb = wl.sh_tables(workload, hrir_dirs=150)
    B, C, order = tb["B"], tb["C"], tb["order"]
    if partitioned:  # 3-partition HRIR (extension a-X)
        hr = tb["hrir_set"]
        hrir = FilterSetMiro.from_arrays(syn.decaying_noise_irs((150, 2, 3 * B - 20), 9, tau=B),
                                         hr._irs_grid.azimuth, hr._irs_grid.colatitude, None, order, True)
        hrir.allow_partitions = True
        hrir.load(B, True, is_prevent_logging=True)
        tb["hrir_set"] = hrir
    conv = wl.make_renderer(tb, S)
    hp_set = _hpcf_set(B, hp_taps, 77)
    conv.set_hpcf(hp_set)
    hnm = tb["hrir_set"].get_filter_blocks_nm()
    cfg = tb["sh_config"]
    cls = oc.ShPartitionedOracle if hnm.shape[0] > 1 else oc.ShOracle
    sh_oracles = [cls(cfg.sh_m, cfg.sh_bases_weighted, hnm, B, order) for _ in range(S)]
    hp_oracles = [oc.OlsOracle(hp_set.get_filter_blocks_fd(), B) for _ in range(S)]
    P_hp = hp_set.get_filter_blocks_fd().shape[0]
    T = 2 * P_hp + 14 if B <= 256 else 6
    x = syn.noise_blocks(T, S * C, B, 5).reshape(T, S, C, B)
    yaw = np.stack([syn.yaw_walk(T, 20 + s, step=20.0) for s in range(S)], axis=1)
    events = {3: ("order", max(order - 2, 0)), 5: ("order", order), 7: ("crossfade", 0), 9: ("crossfade", 1),
              P_hp + 8: ("passthrough", 1), P_hp + 10: ("passthrough", 0), 2 * P_hp + 11: ("clear", 0)}
    if B > 256:
        events = {2: ("crossfade", 0), 4: ("crossfade", 1)}
    worst = _run(conv, sh_oracles, hp_oracles, x, yaw, events)
    print(workload, "S", S, "HPCF partitions", P_hp, "max abs err", worst)
    # detaching restores the plain renderer
    conv.set_hpcf(None)
    y = conv.filter_block(x[0], yaw[0])
    ref = np.stack([o.filter_block(x[0, s], yaw[0, s]) for s, o in enumerate(sh_oracles)])
    assert np.abs(y - ref).max() <= TOL


def test_fused_hpcf_device_path_many_streams():
    """1024 streams on the device entry point: probes against the oracle chain, every launch shape
    of the fused kernel (4 streams per CTA) in use."""
    import torch

    from oracle import sh_chain as oc
    from retisar_b200 import workloads as wl

    tb = wl.sh_tables("c2", hrir_dirs=150)
    S, B, C, order = 1030, tb["B"], tb["C"], tb["order"]
    conv = wl.make_renderer(tb, S)
    hp_set = _hpcf_set(B, 1024, 78)
    conv.set_hpcf(hp_set)
    cfg, hnm = tb["sh_config"], tb["hrir_set"].get_filter_blocks_nm()
    probes = [0, 3, 517, S - 1]
    chains = {s: (oc.ShOracle(cfg.sh_m, cfg.sh_bases_weighted, hnm, B, order),
                  oc.OlsOracle(hp_set.get_filter_blocks_fd(), B)) for s in probes}
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(4)
    idx = torch.as_tensor(probes, device=dev)
    for t in range(8):
        x = torch.randn((S, C, B), generator=gen, device=dev) * 0.1
        yaw = torch.rand((S,), generator=gen, device=dev) * 360.0
        y = conv.filter_block(x, yaw)
        xs, ys, yaws = x[idx].cpu().numpy(), y[idx].cpu().numpy(), yaw[idx].cpu().numpy()
        for i, s in enumerate(probes):
            o, hp = chains[s]
            assert np.abs(ys[i] - hp.filter_block(o.filter_block(xs[i], yaws[i]))).max() <= TOL, (t, s)
