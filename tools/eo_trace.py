"""Phase timeline of one warp of sh_render_eo_kernel (trace build: build.sh -DRTB_TRACE=<cta>,
RTB_LIB_PATH pointing at it).  Prints clock64 deltas per loop phase."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from retisar_b200 import _capi
from retisar_b200 import workloads as wl

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
tb = wl.sh_tables("c2", hrir_dirs=300)
conv = wl.make_renderer(tb, S)
dev = torch.device("cuda", 0)
x = torch.randn((4, S, tb["C"], tb["B"]), device=dev) * 0.1
yaw = torch.rand((4, S), device=dev) * 360
out = torch.empty((S, 2, tb["B"]), device=dev)
for i in range(6):
    conv._plan.process_device(x[i % 4], yaw[i % 4], out, torch.cuda.current_stream())
torch.cuda.synchronize()
buf = np.zeros(1024, dtype=np.int64)
fn = _capi.lib.rtb_debug_eo_trace
fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert fn(buf.ctypes.data, 1024) == 0
t0 = buf[0]
print("prologue -> loop", buf[8] - t0, " loop end", buf[1] - t0, " tail ffts", buf[2] - buf[1], " store", buf[3] - buf[2], " total", buf[3] - t0)
names = ["top", "cpwait", "sync", "readX", "pairsync+pf", "fft", "contract", ""]
for p in range(13):
    b = buf[8 + 8 * p: 8 + 8 * p + 8]
    nxt = buf[8 + 8 * (p + 1)] if p < 12 else buf[1]
    print(p, "cpwait", b[1] - b[0], "sync", b[2] - b[1], "readX", b[3] - b[2], "pair", b[4] - b[3], "fft", b[5] - b[4],
          "contract", b[6] - b[5], "flush", nxt - b[6], "| iter", nxt - b[0])
