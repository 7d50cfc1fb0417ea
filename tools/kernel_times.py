"""Per-kernel device times of the chain WITHOUT flushing L2 between steps (bench.py flushes):
library events on the launching stream.  Usage: kernel_times.py [--workload c2] [--streams S]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from retisar_b200 import _capi
from retisar_b200 import workloads as wl

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2")
ap.add_argument("--streams", type=int, default=1024)
ap.add_argument("--steps", type=int, default=300)
a = ap.parse_args()
tb = wl.sh_tables(a.workload, hrir_dirs=600)
conv = wl.make_renderer(tb, a.streams)
plan = conv._plan
dev = torch.device("cuda", 0)
ring = max(4, int(np.ceil(160e6 / (a.streams * tb["C"] * tb["B"] * 4))) + 1)
x = torch.randn((ring, a.streams, tb["C"], tb["B"]), device=dev) * 0.1
yaw = torch.rand((ring, a.streams), device=dev) * 360
out = torch.empty((a.streams, 2, tb["B"]), device=dev)
st = torch.cuda.current_stream()
for i in range(20):
    plan.process_device(x[i % ring], yaw[i % ring], out, st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for i in range(a.steps):
    plan.process_device(x[i % ring], yaw[i % ring], out, st)
e1.record(st)
torch.cuda.synchronize()
step_us = e0.elapsed_time(e1) / a.steps * 1e3
plan.set_profiling(True)
an, re = [], []
for i in range(a.steps):
    plan.process_device(x[i % ring], yaw[i % ring], out, st)
    an.append(plan.query(_capi.RTB_Q_ANALYSIS_NS))
    re.append(plan.query(_capi.RTB_Q_RENDER_NS))
name = _capi.RENDER_KERNEL_NAMES.get(plan.query(_capi.RTB_Q_RENDER_KERNEL))
print(f"{a.workload} S={a.streams} step {step_us:.1f} us | analysis {np.mean(an) / 1e3:.1f} us, {name} {np.mean(re) / 1e3:.1f} us (events, no flush)")
