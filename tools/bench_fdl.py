"""Measure the FDL kernel (partitioned overlap-save convolver, row a8) on one GPU.

Workload = the reference's own benchmark shape (`_run_benchmark.py`: 2-in/2-out FIR, dirac-like
filter of `--taps` taps, block `--block`), S instances batched in one plan.  Prints one JSON line
with the kernel's HBM roofline (algorithmic bytes / CUDA-event time vs MEASURED_PEAKS.json).
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

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=16384)
ap.add_argument("--taps", type=int, default=4096)
ap.add_argument("--block", type=int, default=256)
ap.add_argument("--channels", type=int, default=2)
ap.add_argument("--mono-in", action="store_true")
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--warmup", type=int, default=20)
a = ap.parse_args()

S, B, C = a.streams, a.block, a.channels
P = -(-a.taps // B)
F = B + 1
Fp = (F + 3) // 4 * 4
n_in = 1 if a.mono_in else C
rng = np.random.default_rng(0)
irs = (rng.standard_normal((C, P * B)) * 0.05 * np.exp(-np.arange(P * B) / (P * B / 6))).astype(np.float32)
H = np.stack([np.fft.rfft(irs[:, p * B:(p + 1) * B], 2 * B) for p in range(P)]).astype(np.complex64)
plan = OlsPlan(S, B, n_in, C, P)
plan.set_filter(H)
dev = torch.device("cuda", 0)
ring = max(2, int(np.ceil(160e6 / (S * n_in * B * 4))) + 1)
x = torch.randn((ring, S, n_in, B), device=dev) * 0.1
out = torch.empty((S, C, B), device=dev)
st = torch.cuda.current_stream()
for i in range(max(a.warmup, P + 2)):
    plan.process_device(x[i % ring], out, st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for i in range(a.steps):
    plan.process_device(x[i % ring], out, st)
e1.record(st)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
plan.set_profiling(True)
ks = []
for i in range(50):
    plan.process_device(x[i % ring], out, st)
    ks.append(plan.query(_capi.RTB_Q_RENDER_NS))
us = float(np.mean(ks)) / 1e3
# algorithmic bytes per stream-block: read P-1 stored spectra + write 1 (per input channel),
# read previous + current input block, write state and output
bytes_per = 8 * Fp * n_in * P + 4 * B * (3 * n_in + C)
peaks = {}
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except OSError:
    pass
peak = float(peaks.get("hbm_gbs", 6650.0))
ach = S * bytes_per / (us * 1e-6) / 1e9
block_rate = 48000 / B
print(json.dumps({
    "metric": "real-time 2-in/2-out partitioned OLS convolvers @48kHz (reference benchmark shape)",
    "value": S / (ms * 1e-3) / block_rate, "unit": "realtime_convolvers", "ms_per_step": ms,
    "config": {"streams": S, "taps": a.taps, "block_length": B, "partitions": P, "n_in": n_in,
               "n_out": C, "input_ring_blocks": ring},
    "roofline": {"bound": "hbm", "kernel": "ols_fdl_kernel", "achieved": ach, "peak": peak,
                 "unit": "GB/s", "frac": ach / peak, "traffic": None,
                 "bytes_per_launch": S * bytes_per, "us_per_launch": us,
                 "peak_source": "measured" if peaks else "fallback"},
}))
