"""Short run of the direction-switching renderer (AdjustableFdConvolver, row f3) for ncu / timing:
S streams, 2 sources, 360 directions, 2048-tap filters at B = 256 (8 partitions)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from retisar_b200.plans import FdPlan

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=4096)
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--taps", type=int, default=2048)
ap.add_argument("--block", type=int, default=256)
a = ap.parse_args()
S, B, NS, ND = a.streams, a.block, 2, 360
P, F = a.taps // B, B + 1
rng = np.random.default_rng(0)
hfd = (rng.standard_normal((P, ND, 2, F)) + 1j * rng.standard_normal((P, ND, 2, F))).astype(np.complex64) * 0.05
plan = FdPlan(S, B, NS, ND, P)
plan.set_filter(hfd)
plan.set_sources(np.array([30.0, -45.0], np.float32), 1)
dev = torch.device("cuda", 0)
x = torch.randn((4, S, NS, B), device=dev) * 0.1
yaw = torch.rand((4, S), device=dev) * 360
out = torch.empty((S, 2, B), device=dev)
st = torch.cuda.current_stream()
for i in range(P + 2):
    plan.process_device(x[i % 4], yaw[i % 4], out, st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for i in range(a.steps):
    plan.process_device(x[i % 4], yaw[i % 4], out, st)
e1.record(st)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
# both output-side queues read + written, input blocks (previous + current) in, state + output out
Fp = (F + 3) // 4 * 4
bytes_per = 2 * 2 * 8 * P * 2 * Fp + 4 * B * (3 * NS + 2)
# the split launch (head kernel + queue kernel) moves less: a consumed head is overwritten by the
# next block instead of cleared and re-read (2P - 2 slot transfers per queue and ear instead of 2P),
# plus the scaled source spectra written by the head kernel and read by the queue kernel
split = os.environ.get("RTB_FD_SPLIT", "1") != "0" and P > 1
moved_per = 2 * 8 * (2 * P - 2) * 2 * Fp + 2 * 8 * NS * Fp + 4 * B * (3 * NS + 2) if split else bytes_per
print(json.dumps({"kernel": "fd_render_kernel + fd_queue_kernel" if split else "fd_render_kernel",
                  "streams": S, "sources": NS, "partitions": P, "block": B,
                  "ms_per_step": ms, "algorithmic_GBps": S * bytes_per / (ms * 1e-3) / 1e9,
                  "bytes_per_launch": S * bytes_per, "moved_GBps": S * moved_per / (ms * 1e-3) / 1e9,
                  "moved_bytes_per_step": S * moved_per, "finite": bool(torch.isfinite(out).all())}))
