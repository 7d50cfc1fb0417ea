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
