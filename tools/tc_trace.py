"""Role timeline of CTA 0 of sh_analysis_tc_kernel (trace build: build.sh -DRTB_TRACE_TC,
RTB_LIB_PATH pointing at it): clock64 stamps of the loader groups, the MMA issuer and the epilogue."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from retisar_b200 import _capi
from retisar_b200 import workloads as wl

name = sys.argv[1] if len(sys.argv) > 1 else "c5"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
tb = wl.sh_tables(name, hrir_dirs=300)
conv = wl.make_renderer(tb, S)
dev = torch.device("cuda", 0)
x = torch.randn((2, S, tb["C"], tb["B"]), device=dev) * 0.1
yaw = torch.rand((2, S), device=dev) * 360
out = torch.empty((S, 2, tb["B"]), device=dev)
for i in range(4):
    conv._plan.process_device(x[i % 2], yaw[i % 2], out, torch.cuda.current_stream())
torch.cuda.synchronize()
buf = np.zeros(4096, dtype=np.int64)
fn = _capi.lib.rtb_debug_tc_trace
fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert fn(buf.ctypes.data, 4096) == 0
t0 = buf[2000]
def rel(v): return int(v - t0)
T = np.arange(20, 60)
def st(role, k, idx=T): return buf[role * 1000 + idx * 8 + k].astype(np.float64)
print("cycles per tile (mma start to start):", np.diff(st(2, 0)).mean())
N = np.arange(40, 99)
print(f"loader warp 0 per chunk: wait_empty {np.mean(st(0,1,N)-st(0,0,N)):.0f}  data wait + split + scatter {np.mean(st(0,2,N)-st(0,1,N)):.0f}  "
      f"load issue {np.mean(st(0,0,N)[1:]-st(0,2,N)[:-1]):.0f}  period {np.diff(st(0,0,N)).mean():.0f}")
print(f"mma per tile: wait acc {np.mean(st(2,1)-st(2,0)):.0f}  wait first chunk {np.mean(st(2,2)-st(2,1)):.0f}  "
      f"chunks issued {np.mean(st(2,6)-st(2,2)):.0f}")
print(f"epilogue per tile: wait tfull {np.mean(st(3,1)-st(3,0)):.0f}  drain {np.mean(st(3,2)-st(3,1)):.0f}")
