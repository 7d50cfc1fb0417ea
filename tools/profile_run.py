"""Short run of the render chain for ncu: a few blocks of S streams (device-resident inputs)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from retisar_b200 import workloads as wl  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2")
ap.add_argument("--streams", type=int, default=1024)
ap.add_argument("--steps", type=int, default=6)
a = ap.parse_args()
tb = wl.sh_tables(a.workload, hrir_dirs=600)
conv = wl.make_renderer(tb, a.streams)
dev = torch.device("cuda", 0)
x = torch.randn((4, a.streams, tb["C"], tb["B"]), device=dev) * 0.1
yaw = torch.rand((4, a.streams), device=dev) * 360
out = torch.empty((a.streams, 2, tb["B"]), device=dev)
for i in range(a.steps):
    conv._plan.process_device(x[i % 4], yaw[i % 4], out, torch.cuda.current_stream())
torch.cuda.synchronize()
print("ok", float(out.abs().max()))
