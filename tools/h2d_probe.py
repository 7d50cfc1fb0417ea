"""Host-to-device copy rate of one c2 input block (33.5 MB) from torch-pinned, library-pinned and
write-combined host memory; decides what bench.py's e2e leg and callers should allocate."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from retisar_b200.plans import PinnedBuffer

n = 1024 * 32 * 256
dev = torch.device("cuda", 0)
dst = torch.empty(n, device=dev)
keep = [PinnedBuffer((n,)), PinnedBuffer((n,), write_combined=True)]
srcs = {"torch pin_memory": torch.randn(n).pin_memory(),
        "rtb_host_alloc": torch.from_numpy(keep[0].array),
        "rtb_host_alloc write-combined": torch.from_numpy(keep[1].array)}
for name, src in srcs.items():
    if "rtb" in name:
        src.numpy()[:] = 0.5
    for _ in range(3):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 50
    print(f"{name:32s} {n * 4 / dt / 1e9:6.1f} GB/s  ({dt * 1e3:.3f} ms per 33.5 MB block)")
