"""Pin the numpy oracle (oracle/sh_chain.py) against fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py).  The restatement performs the same numpy calls, so the bar here is
bit-exact (np.array_equal) except where noted."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import sh_chain as oc

SH_CASES = ["sh_em32_n4_b256", "sh_em32_n4_b64_events", "sh_le110_n8_b128", "sh_rnd50_n3_b32",
            "sh_em32_n2_b64_general", "sh_rnd434_n12_b64"]
OLS_CASES = ["ols_2ch_1024_b256", "ols_2ch_dirac_b128", "ols_hpcf_1024_b64",
             "ols_arir_mono_6ch_b64", "ols_short_b256"]


def run_sh_oracle(g, cls=oc.ShOracle):
    o = cls(g["sh_m"], g["Yw"], g["Hnm"], int(g["B"]), int(g["order"]),
            source_azim_deg=float(g["source_azim"]), is_hrir=bool(g["is_hrir"]))
    ys = []
    for t in range(g["x"].shape[0]):
        kind, val = g["events"][t]
        if kind == 1:
            o.set_order(int(val))
        elif kind == 2:
            o.set_crossfade(bool(val))
        elif kind == 3:
            o.is_passthrough = bool(val)
        elif kind == 4:
            o.clear()
        ys.append(o.filter_block(g["x"][t], g["yaw"][t]).copy())
    return np.stack(ys)


@pytest.mark.parametrize("name", SH_CASES)
def test_sh_oracle_matches_reference(name):
    g = load_golden(name)
    y = run_sh_oracle(g)
    assert y.dtype == np.float32
    assert np.array_equal(y, g["y"]), np.abs(y - g["y"]).max()
    assert np.array_equal(oc.crossfade_windows(int(g["B"]))[0], g["w_in"])
    assert np.array_equal(oc.crossfade_windows(int(g["B"]))[1], g["w_out"])
    assert list(g["rev"]) == oc.reverse_mn_ids(int(g["order"]))


@pytest.mark.parametrize("name", SH_CASES)
def test_partitioned_oracle_reduces_to_reference_for_p1(name):
    """Row a-X is defined by composition; with P == 1 it must reproduce the reference."""
    g = load_golden(name)
    if np.any(g["events"][:, 0] == 3) or np.any(g["events"][:, 0] == 4):
        pytest.skip("passthrough / clear are not part of the a-X definition")
    y = run_sh_oracle(g, oc.ShPartitionedOracle)
    np.testing.assert_allclose(y, g["y"], rtol=0, atol=2e-7)


@pytest.mark.parametrize("name", OLS_CASES)
def test_ols_oracle_matches_reference(name):
    g = load_golden(name)
    B = int(g["B"])
    irs = oc.zero_pad(g["irs"][np.newaxis], B)
    assert np.array_equal(irs, g["irs_td_padded"])
    H = oc.filter_blocks_fd(irs, B)
    assert np.array_equal(H, g["Hfd"])
    o = oc.OlsOracle(H, B)
    for t in range(g["x"].shape[0]):
        kind, val = g["events"][t]
        if kind == 3:
            o.is_passthrough = bool(val)
        elif kind == 4:
            o.clear()
        y = o.filter_block(g["x"][t])
        assert np.array_equal(y, g["y"][t]), (t, np.abs(y - g["y"][t]).max())


def test_ols_oracle_is_a_convolution():
    """Known-answer: partitioned OLS == direct linear convolution."""
    g = load_golden("ols_2ch_1024_b256")
    x = np.concatenate(list(g["x"]), axis=-1).astype(np.float64)
    y = np.concatenate(list(g["y"]), axis=-1)
    for ch in range(2):
        direct = np.convolve(x[ch], g["irs"][ch].astype(np.float64))[: y.shape[-1]]
        np.testing.assert_allclose(y[ch], direct, atol=5e-6)


@pytest.mark.parametrize("name", ["tables_em32_n4", "tables_le110_n8_pinv", "tables_rnd50_n3"])
def test_tables_oracle_matches_reference(name):
    g = load_golden(name)
    w = g["w"] if g["w"].size else None
    sh_m, Yw = oc.sh_configuration(int(g["order"]), g["az"], g["co"], w,
                                   enforce_pinv=bool(g["enforce_pinv"]))
    assert np.array_equal(sh_m, g["sh_m"])
    assert np.array_equal(Yw, g["Yw"])
    B = int(g["B"])
    irs = oc.zero_pad(g["hrir_td"], B)
    assert list(irs.shape) == list(g["hrir_td_padded_shape"])
    Hfd = oc.filter_blocks_fd(irs, B)
    assert np.array_equal(Hfd, g["hrir_blocks_fd"])
    _, hYw = oc.sh_configuration(int(g["order"]), g["hrir_az"], g["hrir_co"], None)
    assert np.array_equal(hYw, g["hrir_Yw"])
    assert np.array_equal(oc.filter_blocks_nm(Hfd, hYw), g["hrir_blocks_nm"])


def test_yaw_chain_matches_reference():
    g = load_golden("yaw_chain")
    m = g["sh_m"]
    with np.errstate(all="ignore"):
        for tag, src, is_hrir in (("hrir", 0.0, True), ("src", float(g["src_azim"]), True),
                                  ("brir", 0.0, False)):
            src16 = np.array([(src, 0)], dtype=np.float16)
            for i, yaw in enumerate(g["yaw"]):
                az = oc.individual_azimuth_deg(yaw, src16, is_hrir)
                assert az.dtype == np.float16
                assert np.array_equal(az[0], g[f"az_{tag}"][i], equal_nan=True)
                ph = oc.sh_rotation_phases(az, m)
                assert np.array_equal(ph, g[f"ph_{tag}"][i], equal_nan=True)


FD_CASES = ["fd_hrir_1src_b64", "fd_brir_2src_b128", "fd_hrir_3src_b32"]


def fd_filter_set(g):
    from retisar_b200 import synthetic as syn

    irs = syn.decaying_noise_irs(tuple(g["irs_shape"]), int(g["irs_seed"]))
    assert abs(np.abs(irs.astype(np.float64)).sum() - float(g["irs_sum"])) < 1e-6, "random stream drifted"
    return irs


@pytest.mark.parametrize("name", FD_CASES)
def test_fd_oracle_matches_reference(name):
    """AdjustableFdConvolver restatement vs the unmodified reference (bit exact)."""
    g = load_golden(name)
    B = int(g["B"])
    H = oc.filter_blocks_fd(oc.zero_pad(fd_filter_set(g), B), B)
    o = oc.FdOracle(H, B, g["sources"], bool(g["is_hrir"]))
    for t in range(g["x"].shape[0]):
        kind, val = g["events"][t]
        if kind == 2:
            o.set_crossfade(bool(val))
        elif kind == 3:
            o.is_passthrough = bool(val)
        elif kind == 4:
            o.clear()
        y = o.filter_block(g["x"][t], g["yaw"][t])
        assert np.array_equal(y, g["y"][t]), (t, np.abs(y - g["y"][t]).max())


def test_oracle_composition_matches_the_references_production_chain():
    """ARIR pre-renderer -> SH renderer -> HPCF run with the reference's own three convolvers
    (`oracle/make_golden.golden_production_chain`): the composition of the oracles reproduces it bit
    for bit, including the pause / clear in the middle."""
    g = load_golden("chain_em32_n4_b64")
    B, order = int(g["B"]), int(g["order"])
    arir = oc.OlsOracle(oc.filter_blocks_fd(oc.zero_pad(g["arir_irs"][None], B), B), B)
    hpcf = oc.OlsOracle(oc.filter_blocks_fd(oc.zero_pad(g["hpcf_irs"][None], B), B), B)
    rend = oc.ShOracle(g["sh_m"], g["Yw"], g["Hnm"], B, order)
    for t in range(g["x"].shape[0]):
        if t == int(g["clear_at"]):
            for o in (arir, rend, hpcf):
                o.clear()
        r = rend.filter_block(arir.filter_block(g["x"][t]), float(g["yaw"][t]))
        assert np.array_equal(r, g["y_renderer"][t]), t
        assert np.array_equal(hpcf.filter_block(r), g["y"][t]), t
