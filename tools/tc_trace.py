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
print("tile | loader g0: wait_empty done-wait arrive | loader g1 ... | mma: start gotacc full0 full1 committed | epi: start gotfull released")
for it in range(20, 26):
    l0 = [rel(buf[0 * 1000 + it * 8 + k]) for k in range(3)]
    l1 = [rel(buf[1 * 1000 + it * 8 + k]) for k in range(3)]
    m = [rel(buf[2 * 1000 + it * 8 + k]) for k in (0, 1, 2, 3, 6)]
    e = [rel(buf[3 * 1000 + it * 8 + k]) for k in range(3)]
    print(it, l0, l1, m, e)
T = np.arange(20, 60)
def st(role, k): return buf[role * 1000 + T * 8 + k].astype(np.float64)
print("cycles per tile (mma start to start):", np.diff(st(2, 0)).mean())
for g in (0, 1):
    print(f"loader g{g}: wait_empty {np.mean(st(g,1)-st(g,0)):.0f}  split+scatter {np.mean(st(g,2)-st(g,1)):.0f}  "
          f"arrive->next wait_empty (load issue + data wait of the other buffer) {np.mean(st(g,0)[1:]-st(g,2)[:-1]):.0f}")
print(f"mma: wait acc {np.mean(st(2,1)-st(2,0)):.0f}  wait full0 {np.mean(st(2,2)-st(2,1)):.0f}  "
      f"chunk0 issue + wait full1 {np.mean(st(2,3)-st(2,2)):.0f}  chunk1 issue {np.mean(st(2,6)-st(2,3)):.0f}")
print(f"epilogue: wait tfull {np.mean(st(3,1)-st(3,0)):.0f}  drain {np.mean(st(3,2)-st(3,1)):.0f}")
