"""BASELINE config c4 as the reference chains it in production: mono source
-> 434-channel ARIR pre-renderer (4096 taps = 64 partitions at 64-sample blocks) -> SH renderer of
order 12 with per-stream yaw -> 1024-tap headphone compensation filter, S batched streams, device
tensors end to end.  Prints one JSON line: block time (CUDA events, inputs cycling through a ring),
real-time streams sustained, and the share of each convolver.

    python tools/bench_c4_chain.py [--streams 1024] [--steps 200] [--warmup 20] [--arir-taps 4096]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from retisar_b200 import workloads as wl

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=1024)
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--warmup", type=int, default=20)
ap.add_argument("--arir-taps", type=int, default=4096)
ap.add_argument("--hpcf-taps", type=int, default=1024)
ap.add_argument("--fused", action="store_true",
                help="ARIR folded into the analysis domain (set_pre_renderer) and HPCF fused into the render "
                     "kernel (set_hpcf): two convolver launches groups instead of three separate convolvers")
a = ap.parse_args()

tb = wl.sh_tables("c4", hrir_dirs=600)
S, B, C = a.streams, tb["B"], tb["C"]
arir, rend, hpcf, arir_irs, _ = wl.make_c4_chain(tb, S, arir_taps=a.arir_taps, hpcf_taps=a.hpcf_taps)
if a.fused:
    rend.set_pre_renderer(arir_irs)
    rend.set_hpcf(hpcf._filter)
    arir = hpcf = None
dev = torch.device("cuda", 0)
ring = 16
x = torch.randn((ring, S, 1, B), device=dev) * 0.5
yaw = torch.rand((ring, S), device=dev) * 360
st = torch.cuda.current_stream()


def step(i, ev=None):
    if a.fused:
        if ev:
            ev[0].record(st); ev[1].record(st)
        y = rend.filter_block(x[i % ring], yaw[i % ring])
        if ev:
            ev[2].record(st); ev[3].record(st)
        return y
    if ev: ev[0].record(st)
    z = arir.filter_block(x[i % ring])
    if ev: ev[1].record(st)
    r = rend.filter_block(z, yaw[i % ring])
    if ev: ev[2].record(st)
    y = hpcf.filter_block(r)
    if ev: ev[3].record(st)
    return y


for i in range(a.warmup):
    step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for i in range(a.steps):
    y = step(i)
e1.record(st)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
parts = np.zeros(3)
n_part = min(a.steps, 50)
for i in range(n_part):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    step(i, ev)
    torch.cuda.synchronize()
    parts += [ev[k].elapsed_time(ev[k + 1]) for k in range(3)]
parts /= n_part
period_ms = B / wl.FS * 1e3
P_arir, P_hpcf = a.arir_taps // B, a.hpcf_taps // B
print(json.dumps({
    "metric": "real-time binaural streams @48kHz/64-blk per GPU (c4 production chain)",
    "value": S * period_ms / ms, "unit": "realtime_streams", "ms_per_step": ms, "block_period_ms": period_ms,
    "config": {"workload": "c4 chain: mono source -> 434-ch ARIR pre-renderer -> SH renderer (order 12, B=64) -> HPCF",
               "fused": bool(a.fused),
               "streams": S, "block_length": B, "channels": C, "arir_taps": a.arir_taps, "arir_partitions": P_arir,
               "hpcf_taps": a.hpcf_taps, "hpcf_partitions": P_hpcf, "input_ring_blocks": ring},
    "ms_per_convolver": {"arir_pre_renderer": float(parts[0]), "sh_renderer": float(parts[1]), "hpcf": float(parts[2])},
    # ARIR stage: mono spectrum history read (P x Fp complex) + C output rows written per stream-block
    "arir_output_GBps": S * C * B * 4 / (parts[0] * 1e-3) / 1e9,
    "finite": bool(torch.isfinite(y).all().item()), "peak": float(y.abs().max().item()),
}))
