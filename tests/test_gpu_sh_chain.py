"""GPU parity tests of the SH render chain, called through the C ABI (ctypes).

Bar (north star): max |y_cuda - y_reference| <= 1e-5 of full scale (1.0) in float32.  The golden
outputs come from the unmodified reference (numpy 2: float32 FFTs), the oracle is bit-identical to
them (tests/test_oracle_vs_golden.py)."""
import numpy as np
import pytest

from conftest import check_regression_bar, load_golden

pytestmark = pytest.mark.gpu

TOL = 1e-5  # absolute, full scale = 1.0

SH_CASES = ["sh_em32_n4_b256", "sh_em32_n4_b64_events", "sh_le110_n8_b128", "sh_rnd50_n3_b32",
            "sh_em32_n2_b64_general", "sh_rnd434_n12_b64"]


def make_plan(g, n_streams=1):
    from retisar_b200.plans import ShPlan

    B, order = int(g["B"]), int(g["order"])
    plan = ShPlan(n_streams, B, g["Yw"].shape[1], order)
    plan.set_tables(g["Yw"], g["Hnm"], g["sh_m"], g["rev"], g["w_in"], g["w_out"])
    plan.set_source(float(g["source_azim"]), 1 if bool(g["is_hrir"]) else -1)
    return plan


def apply_event(plan, kind, val):
    if kind == 1:
        plan.set_order(int(val))
    elif kind == 2:
        plan.set_crossfade(bool(val))
    elif kind == 3:
        plan.set_passthrough(bool(val))
    elif kind == 4:
        plan.reset()


@pytest.mark.parametrize("name", SH_CASES)
def test_sh_chain_matches_reference_golden(name):
    g = load_golden(name)
    plan = make_plan(g)
    worst = 0.0
    for t in range(g["x"].shape[0]):
        apply_event(plan, *g["events"][t])
        y = plan.process_host(g["x"][t][None], g["yaw"][t:t + 1])[0]
        err = np.abs(y - g["y"][t]).max()
        worst = max(worst, err)
        assert err <= TOL, (name, t, err)
    check_regression_bar(float(worst), float(np.abs(g["y"]).max()), name)


def test_symmetric_mode_detected():
    from retisar_b200 import _capi

    g = load_golden("sh_em32_n4_b256")
    plan = make_plan(g)
    assert plan.query(_capi.RTB_Q_SYMMETRIC) == 1
    assert plan.query(_capi.RTB_Q_ROWS) == 25
    assert plan.query(_capi.RTB_Q_SIGNALS) == 15
    g = load_golden("sh_em32_n2_b64_general")
    plan = make_plan(g)
    assert plan.query(_capi.RTB_Q_SYMMETRIC) == 0
    assert plan.query(_capi.RTB_Q_ROWS) == 18


@pytest.mark.parametrize("name", ["sh_em32_n4_b256", "sh_em32_n4_b64_events"])
def test_batched_streams_are_independent(name):
    """S streams with permuted inputs / yaw offsets reproduce S oracle runs."""
    from oracle import sh_chain as oc

    g = load_golden(name)
    S, T = 7, g["x"].shape[0]
    B, order = int(g["B"]), int(g["order"])
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((T, S, g["Yw"].shape[1], B)) * 0.1).astype(np.float32)
    yaw = rng.uniform(-400, 400, (T, S)).astype(np.float32)
    plan = make_plan(g, S)
    oracles = [oc.ShOracle(g["sh_m"], g["Yw"], g["Hnm"], B, order) for _ in range(S)]
    for t in range(T):
        kind, val = g["events"][t]
        apply_event(plan, kind, val)
        for o in oracles:
            if kind == 1:
                o.set_order(int(val))
            elif kind == 2:
                o.set_crossfade(bool(val))
            elif kind == 3:
                o.is_passthrough = bool(val)
            elif kind == 4:
                o.clear()
        y = plan.process_host(x[t], yaw[t])
        ref = np.stack([o.filter_block(x[t, s], yaw[t, s]) for s, o in enumerate(oracles)])
        assert np.abs(y - ref).max() <= TOL, (t, np.abs(y - ref).max())
