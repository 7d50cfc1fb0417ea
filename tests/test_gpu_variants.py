"""The kernel variants that are off by default (fallbacks selected by geometry or environment) stay
parity-green: each is switched on through its environment variable before the
plan is created and run against a reference golden with control events."""
import numpy as np
import pytest

from conftest import load_golden
from test_gpu_sh_chain import apply_event, make_plan

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.mark.parametrize("env", [
    {"RTB_RENDER_MG": "0"},            # term-wise render kernel only
    {"RTB_PDL": "1"},                  # programmatic dependent launch
    {"RTB_ANALYSIS_TC": "0"},          # mma.sync (warp-level tensor core) analysis GEMM
    {"RTB_ANALYSIS_TC": "0", "RTB_ANALYSIS_MMA": "0"},  # FFMA2 SIMT analysis GEMM
    {"RTB_TC_NST": "2"},               # tcgen05 analysis with two operand stages (4-warp loader groups)
])
def test_variant_matches_reference_golden(env, monkeypatch):
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    g = load_golden("sh_em32_n4_b256")
    S = 9
    plan = make_plan(g, S)
    for t in range(g["x"].shape[0]):
        apply_event(plan, *g["events"][t])
        x = np.repeat(g["x"][t][None], S, axis=0)
        yaw = np.repeat(g["yaw"][t:t + 1], S)
        y = plan.process_host(x, yaw)
        assert np.abs(y - g["y"][t][None]).max() <= TOL, (env, t)
    # device entry point
    import torch

    dev = torch.device("cuda", 0)
    plan2 = make_plan(g, S)
    out = torch.empty((S, 2, int(g["B"])), device=dev)
    for t in range(g["x"].shape[0]):
        apply_event(plan2, *g["events"][t])
        x = torch.from_numpy(np.repeat(g["x"][t][None], S, axis=0)).to(dev)
        yaw = torch.from_numpy(np.repeat(g["yaw"][t:t + 1], S)).to(dev)
        plan2.process_device(x, yaw, out, torch.cuda.current_stream())
        torch.cuda.synchronize()
        assert np.abs(out.cpu().numpy() - g["y"][t][None]).max() <= TOL, (env, "device", t)


@pytest.mark.parametrize("name", ["sh_em32_n4_b64_events", "sh_rnd50_n3_b32", "sh_em32_n2_b64_general",
                                  "sh_rnd434_n12_b64"])
def test_m_grouped_kernel_on_small_blocks(name, monkeypatch):
    """Below 128-sample blocks several streams share a warp in the m-grouped render kernel (it is the
    default only for batches of 2 048 streams and more): forced here on the small-block goldens of the
    unmodified reference, with more streams than one CTA holds and a ragged last warp."""
    from retisar_b200 import _capi

    monkeypatch.setenv("RTB_RENDER_MG", "1")
    g = load_golden(name)
    S = 21
    plan = make_plan(g, S)
    used = set()
    for t in range(g["x"].shape[0]):
        apply_event(plan, *g["events"][t])
        x = np.repeat(g["x"][t][None], S, axis=0)
        yaw = np.repeat(g["yaw"][t:t + 1], S)
        y = plan.process_host(x, yaw)
        used.add(_capi.RENDER_KERNEL_NAMES[plan.query(_capi.RTB_Q_RENDER_KERNEL)])
        assert np.abs(y - g["y"][t][None]).max() <= TOL, (name, t)
    assert "sh_render_mg_kernel" in used, used

