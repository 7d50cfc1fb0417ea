"""BASELINE config c4 as the reference chains it in production: mono source -> ARIR pre-renderer
(434-channel partitioned OLS, mono input broadcast) -> SH renderer (order 12, 64-sample blocks,
per-stream yaw) -> headphone compensation filter (2-channel partitioned OLS).  Device tensors all
the way; checked block by block against the composition of the pinned oracles."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-5


def test_c4_production_chain_matches_oracle_composition():
    import torch

    from oracle import sh_chain as oc
    from retisar_b200 import synthetic as syn
    from retisar_b200 import workloads as wl

    S, T = 3, 24
    tb = wl.sh_tables("c4", hrir_dirs=200)
    B, C = tb["B"], tb["C"]
    arir, rend, hpcf, arir_irs, hpcf_irs = wl.make_c4_chain(tb, S, arir_taps=512, hpcf_taps=1024)
    cfg = tb["sh_config"]
    hnm = rend._filter.get_filter_blocks_nm()
    o_arir = [oc.OlsOracle(oc.filter_blocks_fd(arir_irs[None], B), B) for _ in range(S)]
    o_rend = [oc.ShOracle(cfg.sh_m, cfg.sh_bases_weighted, hnm, B, tb["order"]) for _ in range(S)]
    o_hpcf = [oc.OlsOracle(oc.filter_blocks_fd(hpcf_irs[None], B), B) for _ in range(S)]

    dev = torch.device("cuda", 0)
    x = syn.noise_blocks(T, S, B, 11, scale=0.5).reshape(T, S, 1, B)
    yaw = np.stack([syn.yaw_walk(T, 20 + s) for s in range(S)], axis=1).astype(np.float32)
    worst = peak = 0.0
    for t in range(T):
        a = arir.filter_block(torch.from_numpy(x[t]).to(dev))            # [S, 434, B]
        assert a.shape == (S, C, B)
        r = rend.filter_block(a, torch.from_numpy(yaw[t]).to(dev))      # [S, 2, B]
        y = hpcf.filter_block(r).cpu().numpy()                          # [S, 2, B]
        for s in range(S):
            ref = o_hpcf[s].filter_block(o_rend[s].filter_block(o_arir[s].filter_block(x[t, s]), yaw[t, s]))
            worst = max(worst, float(np.abs(y[s] - ref).max()))
            peak = max(peak, float(np.abs(ref).max()))
    assert peak > 1e-3  # the chain carries signal
    assert worst <= TOL, worst


@pytest.mark.parametrize("workload,S,arir_taps,hp_taps", [("c4", 5, 512, 1024), ("c2", 3, 1000, 0), ("c4", 130, 320, 256)])
def test_chain_with_folded_pre_renderer_and_fused_hpcf(workload, S, arir_taps, hp_taps):
    """The same production chain with the ARIR folded into the analysis domain
    (`AdjustableShConvolver.set_pre_renderer`: the pre-renderer writes the analysis rows, no
    434-channel intermediate, no analysis GEMM) and the HPCF fused into the render kernel
    (`set_hpcf`), against the composition of the three pinned oracles on the ORIGINAL filters."""
    import torch

    from oracle import sh_chain as oc
    from retisar_b200 import FilterSet
    from retisar_b200 import synthetic as syn
    from retisar_b200 import workloads as wl

    tb = wl.sh_tables(workload, hrir_dirs=150)
    B, C, order = tb["B"], tb["C"], tb["order"]
    arir_irs = syn.decaying_noise_irs((C, arir_taps), 40)
    rend = wl.make_renderer(tb, S)
    rend.set_pre_renderer(arir_irs)
    J = rend._plan.query(2)
    assert rend._pre_renderer._plan.n_out == J and rend._pre_renderer._plan.n_in == 1
    o_hpcf = None
    if hp_taps:
        hpcf_irs = syn.decaying_noise_irs((2, hp_taps), 41, scale=0.3)
        hp = FilterSet.create_instance_by_type(hpcf_irs, "FIR_MULTICHANNEL")
        hp.load(block_length=B, is_single_precision=True, is_prevent_logging=True)
        rend.set_hpcf(hp)
        o_hpcf = {}
    cfg, hnm = tb["sh_config"], rend._filter.get_filter_blocks_nm()
    probes = sorted({0, S // 2, S - 1})
    o_arir = {s: oc.OlsOracle(oc.filter_blocks_fd(oc.zero_pad(arir_irs[None], B), B), B) for s in probes}
    o_rend = {s: oc.ShOracle(cfg.sh_m, cfg.sh_bases_weighted, hnm, B, order) for s in probes}
    if hp_taps:
        o_hpcf = {s: oc.OlsOracle(oc.filter_blocks_fd(oc.zero_pad(hpcf_irs[None], B), B), B) for s in probes}
    dev = torch.device("cuda", 0)
    T = arir_taps // B + 8
    x = syn.noise_blocks(T, S, B, 12, scale=0.5).reshape(T, S, 1, B)
    yaw = np.stack([syn.yaw_walk(T, 60 + s, step=15.0) for s in range(S)], axis=1).astype(np.float32)
    worst = peak = 0.0
    for t in range(T):
        if t == T - 3:  # pause protocol: every client clears its buffers
            rend._clear_buffers()
            for d in (o_arir, o_rend) + ((o_hpcf,) if o_hpcf else ()):
                for o in d.values():
                    o.clear()
        y = rend.filter_block(torch.from_numpy(x[t]).to(dev), torch.from_numpy(yaw[t]).to(dev)).cpu().numpy()
        for s in probes:
            ref = o_rend[s].filter_block(o_arir[s].filter_block(x[t, s]), yaw[t, s])
            if o_hpcf:
                ref = o_hpcf[s].filter_block(ref)
            worst = max(worst, float(np.abs(y[s] - ref).max()))
            peak = max(peak, float(np.abs(ref).max()))
    print(workload, "S", S, "rows", J, "max abs err", worst, "peak", peak)
    assert peak > 1e-3 and worst <= TOL, (worst, peak)
    # numpy in, numpy out
    y = rend.filter_block(x[0], yaw[0])
    assert isinstance(y, np.ndarray) and y.shape == (S, 2, B)


def test_folded_chain_matches_the_references_own_three_convolvers():
    """Golden `chain_em32_n4_b64`: mono source -> ARIR pre-renderer -> SH renderer -> HPCF rendered by
    the UNMODIFIED reference's three convolver objects (`oracle/make_golden.golden_production_chain`).
    Here: one `AdjustableShConvolver` with the pre-renderer folded into the analysis domain and the
    HPCF fused into the render kernel, 5 identical streams, pause / clear in the middle."""
    import torch

    from conftest import load_golden
    from retisar_b200 import Convolver, FilterSet
    from retisar_b200 import synthetic as syn
    from retisar_b200._filter_set import FilterSetMiro

    g = load_golden("chain_em32_n4_b64")
    B, order, S = int(g["B"]), int(g["order"]), 5
    az, co, w = syn.load_array_grid("em32")
    arr = FilterSetMiro.from_arrays(np.zeros((1, len(az), B), np.float32), az, co, w, order, False)
    arr.load(B, True, is_prevent_logging=True)
    haz, hco = syn.random_sphere_grid(120, 701)   # as oracle/make_golden._array_and_hrir(seed = 700)
    hrir = FilterSetMiro.from_arrays(syn.decaying_noise_irs((120, 2, 64), 702), haz, hco, None, order, True)
    hrir.load(B, True, is_prevent_logging=True)
    rend = Convolver.create_instance_by_filter_set(hrir, B, [(0, 0)], None, n_streams=S)
    rend.prepare_sh_processing(arr.get_sh_configuration(), 0, None,
                               comp_nm=np.ones(((order + 1) ** 2, 1, B + 1), np.complex64))
    # same tables as the reference built (to the last bits: the host BLAS of the box may order the sums differently)
    assert np.abs(rend._filter.get_filter_blocks_nm() - g["Hnm"]).max() <= 1e-6 * np.abs(g["Hnm"]).max()
    rend.set_pre_renderer(g["arir_irs"])
    hp = FilterSet.create_instance_by_type(g["hpcf_irs"], "FIR_MULTICHANNEL")
    hp.load(block_length=B, is_single_precision=True, is_prevent_logging=True)
    rend.set_hpcf(hp)
    dev = torch.device("cuda", 0)
    worst = 0.0
    for t in range(g["x"].shape[0]):
        if t == int(g["clear_at"]):
            rend._clear_buffers()
        x = torch.from_numpy(np.repeat(g["x"][t][None], S, axis=0)).to(dev)
        yaw = torch.full((S,), float(g["yaw"][t]), device=dev)
        y = rend.filter_block(x, yaw).cpu().numpy()
        err = float(np.abs(y - g["y"][t][None]).max())
        worst = max(worst, err)
        assert err <= TOL, (t, err)
    print("folded chain vs the reference's three convolvers: max abs err", worst, "peak", np.abs(g["y"]).max())
