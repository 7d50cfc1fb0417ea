"""Timeline of CTA 0 of sh_analysis_tc_kernel for a SMALL batch (every CTA has only a dozen tiles):
clock64 stamps of loader group 0..3, the MMA issuer and the epilogue per tile, relative to the first
stamp.  Trace build: build.sh -DRTB_TRACE_TC, RTB_LIB_PATH pointing at it."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from retisar_b200 import _capi
from retisar_b200 import workloads as wl

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
tb = wl.sh_tables(name, hrir_dirs=300)
conv = wl.make_renderer(tb, S)
dev = torch.device("cuda", 0)
x = torch.randn((8, S, tb["C"], tb["B"]), device=dev) * 0.1
yaw = torch.rand((8, S), device=dev) * 360
out = torch.empty((S, 2, tb["B"]), device=dev)
flush = torch.empty(int(256e6) // 4, device=dev)
for i in range(6):
    flush.zero_()
    conv._plan.process_device(x[i % 8], yaw[i % 8], out, torch.cuda.current_stream())
torch.cuda.synchronize()
buf = np.zeros(4096, dtype=np.int64)
fn = _capi.lib.rtb_debug_tc_trace
fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert fn(buf.ctypes.data, 4096) == 0
n_tiles = (S * tb["B"] // 128 + 147) // 148
vals = buf[buf > 0]
t0 = vals.min()
def g(role, it, k): return int(buf[role * 1000 + it * 8 + k] - t0) if buf[role * 1000 + it * 8 + k] > 0 else -1
print("tile | loader: loads issued, stage free, scattered | mma: start, acc free, data ready, committed | epilogue: start, acc full, drained  (cycles from the first stamp)")
for it in range(n_tiles + 1):
    print(it, "|", g(0, it, 0), g(0, it, 1), g(0, it, 2), "|", g(2, it, 0), g(2, it, 1), g(2, it, 2), g(2, it, 6), "|", g(3, it, 0), g(3, it, 1), g(3, it, 2))
print("last stamp", int(vals.max() - t0), "cycles")
