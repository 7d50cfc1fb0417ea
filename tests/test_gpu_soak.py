"""Soak checks: two independent renderers fed the same long sequence of blocks must stay bit-identical
(a race anywhere in the chain shows up as a divergence), and a probe stream must stay on the oracle
(no slow drift of the overlap-save / queue / delay-line state)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("workload,S,blocks", [("c2", 1024, 1500), ("c3", 300, 300), ("c4", 2100, 600)])
def test_twin_renderers_stay_bit_identical_and_on_the_oracle(workload, S, blocks):
    import torch

    from oracle import sh_chain as oc
    from retisar_b200 import workloads as wl

    tb = wl.sh_tables(workload, hrir_dirs=200)
    C, B = tb["C"], tb["B"]
    r1, r2 = wl.make_renderer(tb, S), wl.make_renderer(tb, S)
    cfg, hnm = tb["sh_config"], tb["hrir_set"].get_filter_blocks_nm()
    cls = oc.ShOracle if hnm.shape[0] == 1 else oc.ShPartitionedOracle
    probe = S // 3
    o = cls(cfg.sh_m, cfg.sh_bases_weighted, hnm, B, tb["order"])
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1)
    ring = 16
    x = torch.randn((ring, S, C, B), generator=gen, device=dev) * 0.1
    yaw = torch.rand((ring, S), generator=gen, device=dev) * 360
    xh, yh = x[:, probe].cpu().numpy(), yaw[:, probe].cpu().numpy()
    worst = 0.0
    for t in range(blocks):
        y1 = r1.filter_block(x[t % ring], yaw[t % ring])
        y2 = r2.filter_block(x[t % ring], yaw[t % ring])
        ref = o.filter_block(xh[t % ring], yh[t % ring])
        if t % 50 == 0 or t == blocks - 1:
            assert torch.equal(y1, y2), f"renderers diverged at block {t}"
            worst = max(worst, float(np.abs(y1[probe].cpu().numpy() - ref).max()))
            assert worst <= 1e-5, (t, worst)
    print(f"soak: {workload} x {S} streams x {blocks} blocks, probe max abs err {worst:.3g}")
