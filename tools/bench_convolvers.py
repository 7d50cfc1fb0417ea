"""The reference's own benchmark grid on the B200 FDL kernel.

`ReTiSAR/_run_benchmark.py:38-46, 252-344` keeps adding 2-in/2-out `OverlapSaveConvolver`s with a
dirac filter of {128, 512, 4096, 110250} taps at block lengths 128 ... 4096 until JACK reports
dropouts; the published medians (iMac Pro 2017, `res/research/benchmarks/2020-09-02 iM`, decoded in
SURVEY.md section 6a) are the only throughput numbers the reference ships.  Here the same shapes
run as S batched convolvers in one `rtb_ols_plan`; a configuration "holds" S convolvers if one
block of all S takes less than the block period, so the sustained count is S * period / t_block
(measured at an S that fills the GPU, device-resident inputs from a ring larger than L2).

Prints one JSON line per (taps, block) and a markdown table at the end.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from retisar_b200 import _capi  # noqa: E402
from retisar_b200.plans import OlsPlan  # noqa: E402

# max convolvers, block 128 / 256 / 512 / 1024 / 2048 / 4096
PUBLISHED = {
    128: [1, 142, 357, 585, 813, 762],
    512: [1, 39, 325, 558, 820, 760],
    4096: [4, 8, 71, 187, 343, 755],
    110250: [0, 1, 3, 12, 30, 53],
}
BLOCKS = [128, 256, 512, 1024, 2048, 4096]

ap = argparse.ArgumentParser()
ap.add_argument("--taps", type=int, nargs="*", default=[128, 512, 4096, 110250])
ap.add_argument("--blocks", type=int, nargs="*", default=BLOCKS)
ap.add_argument("--state-gb", type=float, default=12.0, help="cap on the plans' delay-line memory")
ap.add_argument("--steps", type=int, default=30)
a = ap.parse_args()

peaks = {}
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except OSError:
    pass
peak = float(peaks.get("hbm_gbs", 6650.0))
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream()
rows = []
for taps in a.taps:
    for B in a.blocks:
        P = -(-taps // B)
        F = B + 1
        Fp = (F + 3) // 4 * 4
        per_stream_state = 8 * Fp * 2 * P + 4 * 2 * 2 * B
        # enough convolvers to fill the GPU (>= 8 CTA waves worth of channel pairs), capped by memory
        S = int(min(max(1024, 2_000_000 // B), a.state_gb * 1e9 // per_stream_state))
        S = max(S, 8)
        H = np.zeros((P, 2, F), np.complex64)
        H[0] = 1.0  # dirac at tap 0: spectrum of partition 0 is all ones (as `get_dirac_blocks_fd`)
        plan = OlsPlan(S, B, 2, 2, P)
        plan.set_filter(H)
        ring = max(2, int(np.ceil(160e6 / (S * 2 * B * 4))) + 1)
        x = torch.randn((ring, S, 2, B), device=dev) * 0.1
        out = torch.empty((S, 2, B), device=dev)
        warm = P + 2  # every slot of the delay line holds a spectrum before timing starts
        for i in range(warm):
            plan.process_device(x[i % ring], out, st)
        torch.cuda.synchronize()
        # identity check: a dirac filter returns its input
        assert float((out - x[(warm - 1) % ring]).abs().max()) < 1e-5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for i in range(a.steps):
            plan.process_device(x[i % ring], out, st)
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        period_ms = B / 48000 * 1e3
        sustained = S * period_ms / ms
        bytes_per = 8 * Fp * 2 * P + 4 * B * (3 * 2 + 2)
        ach = S * bytes_per / (ms * 1e-3) / 1e9
        pub = PUBLISHED.get(taps, [None] * 6)[BLOCKS.index(B)] if B in BLOCKS else None
        row = {"taps": taps, "block_length": B, "partitions": P, "convolvers_in_batch": S,
               "ms_per_block": ms, "block_period_ms": period_ms, "realtime_convolvers": sustained,
               "reference_published": pub, "hbm_GBps": ach, "hbm_frac": ach / peak}
        rows.append(row)
        print(json.dumps(row), flush=True)
        plan.close()
        del x, out
        torch.cuda.empty_cache()

print("\n| taps | block | partitions | batch | ms / block | real-time convolvers (B200) | reference (iMac Pro, published) | HBM GB/s | of measured peak |")
print("|---|---|---|---|---|---|---|---|---|")
for r in rows:
    print(f"| {r['taps']} | {r['block_length']} | {r['partitions']} | {r['convolvers_in_batch']} | {r['ms_per_block']:.3f} | "
          f"{r['realtime_convolvers']:.0f} | {r['reference_published']} | {r['hbm_GBps']:.0f} | {100 * r['hbm_frac']:.0f} % |")
