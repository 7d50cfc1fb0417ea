"""Block feeder (row f4): the receive / process / deliver cycle of `JackClient`
(`_jack_client.py:130-152, 713-867`) around a convolver.  Level metering is pinned to a fixture
recorded from the reference's own `tools.calculate_rms` / `calculate_peak`."""
import numpy as np
import pytest

from conftest import load_golden


def test_levels_oracle_matches_reference():
    from oracle import sh_chain as oc

    g = load_golden("levels_4ch_b256")
    for t in range(g["x"].shape[0]):
        y = g["x"][t].copy()
        rms, peak = oc.deliver(y, float(g["volume"]), g["relative"])
        assert np.array_equal(y, g["y"][t])
        assert np.array_equal(rms, g["rms"][t]) and np.array_equal(peak, g["peak"][t])
    assert g["rms"][2, 1] == -200 and g["peak"][2, 1] == -200  # silent channel


def test_feeder_host_passthrough_levels_delay_and_dropouts():
    from retisar_b200._feeder import BlockFeeder

    g = load_golden("levels_4ch_b256")
    f = BlockFeeder(None, 256, 48000, n_output_channels=4, output_volume_dbfs=-3.0,
                    is_measure_levels=True)
    for port, rel in enumerate(g["relative"][:, 0]):
        f.set_output_volume_port_relative_db(port, 20 * np.log10(rel))
    assert f.set_output_volume_port_relative_db(7, 0.0) is None  # out of range: ignored
    for t in range(g["x"].shape[0]):
        y = f.process_block(g["x"][t].copy())
        assert np.allclose(y, g["y"][t], rtol=1e-6, atol=0)
        assert np.allclose(f.last_rms, g["rms"][t], atol=1e-4)
        assert np.allclose(f.last_peak, g["peak"][t], atol=1e-4)
    assert f._counter_clipping >= 1          # the x9 channel exceeds full scale
    assert f.set_output_mute(True) is True and f.process_block(g["x"][0].copy()) is None
    f.set_output_mute(False)
    # input delay of two blocks
    f.init_input_buffer(0.1, 4)
    assert f.set_input_delay_ms(10.0) == pytest.approx(2 * 256 / 48000 * 1000)
    f.set_output_volume_db(0.0)
    for port in range(4):
        f.set_output_volume_port_relative_db(port, 0.0)
    blocks = [np.full((4, 256), float(i + 1), np.float32) for i in range(5)]
    outs = [np.array(f.process_block(b.copy())) for b in blocks]
    assert np.all(outs[0] == 0) and np.all(outs[1] == 0) and np.all(outs[2] == 1) and np.all(outs[4] == 3)
    # dropout accounting: a block that takes longer than its period counts once
    f.get_dropout_counter()
    f._block_period = 0.0
    f.process_block(blocks[0].copy())
    assert f.get_dropout_counter() == 1 and f.get_dropout_counter() == 0


@pytest.mark.gpu
def test_feeder_device_output_stage_and_renderer():
    import torch

    from oracle import sh_chain as oc
    from retisar_b200 import workloads as wl
    from retisar_b200._feeder import BlockFeeder

    tb = wl.sh_tables("c2", hrir_dirs=200)
    S, C, B = 37, tb["C"], tb["B"]
    conv = wl.make_renderer(tb, S)
    f = BlockFeeder(conv, B, 48000, output_volume_dbfs=-6.0, is_measure_levels=True)
    f.set_output_volume_port_relative_db(1, 3.0)
    cfg = tb["sh_config"]
    oracles = [oc.ShOracle(cfg.sh_m, cfg.sh_bases_weighted, tb["hrir_set"].get_filter_blocks_nm(), B, tb["order"])
               for _ in range(3)]
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(3)
    vol = 10 ** (-6.0 / 20.0)
    rel = np.array([[1.0], [10 ** (3.0 / 20.0)]])
    for t in range(3):
        x = (rng.standard_normal((S, C, B)) * 0.1).astype(np.float32)
        if t == 1:
            x[4] *= 300.0  # drive one stream into clipping
        yaw = rng.uniform(0, 360, S).astype(np.float32)
        y = f.process_block(torch.from_numpy(x).to(dev), torch.from_numpy(yaw).to(dev))
        yh = y.cpu().numpy()
        for s, o in enumerate(oracles):
            ref = o.filter_block(x[s], yaw[s]).astype(np.float32)
            rms, peak = oc.deliver(ref, vol, rel)
            assert np.abs(yh[s] - ref).max() <= 1e-5
            assert np.allclose(f.last_rms[s], rms, atol=1e-3) and np.allclose(f.last_peak[s], peak, atol=1e-3)
    assert f.get_clipping_counter() >= 1
    assert f.last_rms.shape == (S, 2)


@pytest.mark.gpu
def test_feeder_device_clipping_accounting():
    """Device path: nothing is counted while detection is off; with detection on and level
    metering off the count is still delivered by `get_clipping_counter`, and it ADDS to what the
    host path counted before."""
    import torch

    from retisar_b200._feeder import BlockFeeder

    B = 64
    f = BlockFeeder(None, B, 48000, is_measure_levels=False, is_detect_clipping=False)
    loud = torch.full((5, 2, B), 3.0, device="cuda")
    f.process_block(loud.clone())
    assert f.get_clipping_counter() == 0
    f._is_detect_clipping = True
    f.process_block(np.full((2, B), 2.0, np.float32))      # host path: 2 clipped channel blocks
    assert f.get_clipping_counter() == 2
    f.process_block(loud.clone())                           # device path: 10 more
    assert f.get_clipping_counter() == 12
    f.process_block(loud.clone() * 0.1)                     # no clipping
    assert f.get_clipping_counter() == 12
