"""DelayBuffer (row f4; reference `_convolver.py:11-141`): the oracle and the host class against a
fixture recorded from the unmodified reference (delay changes while running, the maximum-delay
wrap, `None` input), and the batched device ring on the GPU.  Bar: bit-exact (it only moves data)."""
import numpy as np
import pytest

from conftest import load_golden


def _run(buf, g):
    ys = []
    for t in range(g["x"].shape[0]):
        if g["set_delay_ms"][t] >= 0:
            assert buf.set_delay(float(g["set_delay_ms"][t])) == pytest.approx(float(g["realized_ms"][t]), abs=0)
        ys.append(np.array(buf.process_block(g["x"][t].copy())))
    return np.stack(ys)


def test_delay_oracle_matches_reference():
    from oracle import sh_chain as oc

    g = load_golden("delay_3ch_b64")
    buf = oc.DelayBufferOracle(int(g["fs"]), int(g["B"]), float(g["start_delay_ms"]))
    assert buf.set_delay(5.0) == buf.delay_blocks  # before init: the delay in blocks (`:100-101`)
    buf = oc.DelayBufferOracle(int(g["fs"]), int(g["B"]), float(g["start_delay_ms"]))
    buf.init(float(g["max_sec"]), g["x"].shape[1])
    assert buf.buffer.shape[0] == int(g["ring_blocks"])
    assert np.array_equal(_run(buf, g), g["y"])
    assert buf.process_block(None) is None


def test_delay_buffer_class_matches_reference():
    from retisar_b200._convolver import DelayBuffer

    g = load_golden("delay_3ch_b64")
    buf = DelayBuffer(int(g["fs"]), int(g["B"]), float(g["start_delay_ms"]))
    x0 = g["x"][0]
    assert buf.process_block(x0) is x0  # not initialised: passes the block through
    buf.init(float(g["max_sec"]), g["x"].shape[1])
    assert buf._buffer.shape == (int(g["ring_blocks"]), g["x"].shape[1], int(g["B"]))
    assert np.array_equal(_run(buf, g), g["y"])
    assert buf.process_block(None) is None


@pytest.mark.gpu
def test_delay_buffer_batched_on_device():
    import torch

    from oracle import sh_chain as oc
    from retisar_b200._convolver import DelayBuffer

    fs, B, C, S, T = 48000, 128, 4, 5, 25
    dev = torch.device("cuda", 0)
    buf = DelayBuffer(fs, B, delay_ms=8.0)
    buf.init(0.05, C, n_streams=S, device=dev)
    oracles = [oc.DelayBufferOracle(fs, B, 8.0) for _ in range(S)]
    for o in oracles:
        o.init(0.05, C)
    rng = np.random.default_rng(1)
    for t in range(T):
        if t in (6, 11, 17):
            ms = {6: 0.0, 11: 21.0, 17: 500.0}[t]
            r = buf.set_delay(ms)
            assert all(o.set_delay(ms) == r for o in oracles)
        x = rng.standard_normal((S, C, B)).astype(np.float32)
        y = buf.process_block(torch.from_numpy(x).to(dev)).cpu().numpy()
        ref = np.stack([np.array(o.process_block(x[s].copy())) for s, o in enumerate(oracles)])
        assert np.array_equal(y, ref), t
