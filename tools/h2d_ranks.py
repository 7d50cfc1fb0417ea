"""Host <-> device copy rates of N ranks at once (one process per GPU, torchrun): what limits the
end-to-end (host-buffer) scaling of the render chain -- the PCIe link of each GPU, a shared PCIe
switch, or the host memory system.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/h2d_ranks.py [--bind]

Every rank copies one c2 input block (1 024 streams x 32 ch x 256 samples = 33.5 MB) host -> device
and one result block (2 MB) device -> host, 100 times: first each rank ALONE (the others wait at a
barrier), then ALL ranks together.  Rank 0 prints one JSON line with the per-rank GB/s of both
phases plus the topology the driver reports (`nvidia-smi topo -m`).  --bind pins every rank to the
CPUs `nvidia-smi topo -m` lists for its GPU before the pinned buffers are allocated.
"""
import argparse
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def topo():
    try:
        return subprocess.run(["nvidia-smi", "topo", "-m"], stdout=subprocess.PIPE, text=True, timeout=20).stdout
    except Exception as e:  # noqa: BLE001
        return f"unavailable: {e}"


def cpu_affinity_from_topo(text, gpu):
    """CPU list ("0-31,64-95") nvidia-smi reports for GPU<gpu>, or None."""
    for line in text.splitlines():
        cols = line.split()
        if cols and cols[0] == f"GPU{gpu}":
            for c in cols[1:]:
                if c[0].isdigit() and ("-" in c or "," in c) and not c.startswith("NV"):
                    return c
    return None


def parse_cpus(s):
    out = set()
    for part in s.split(","):
        lo, _, hi = part.partition("-")
        out.update(range(int(lo), int(hi or lo) + 1))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bind", action="store_true")
    ap.add_argument("--reps", type=int, default=100)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    t = topo()
    bound = None
    if a.bind:
        cpus = cpu_affinity_from_topo(t, local)
        if cpus:
            want = parse_cpus(cpus) & os.sched_getaffinity(0)
            if want:
                os.sched_setaffinity(0, want)
                bound = cpus
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_in, n_out = 1024 * 32 * 256, 1024 * 2 * 256
    h_in = torch.randn(n_in).pin_memory()
    h_out = torch.empty(n_out).pin_memory()
    d_in, d_out = torch.empty(n_in, device=dev), torch.randn(n_out, device=dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            with torch.cuda.stream(s_in):
                d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s_out):
                h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        return n_in * 4 / dt / 1e9, n_out * 4 / dt / 1e9, dt * 1e3

    def barrier():
        if world > 1:
            dist.barrier()

    run(10)
    alone = [None] * world
    for r in range(world):
        barrier()
        if r == rank:
            alone[r] = run(a.reps)
        barrier()
    barrier()
    together = run(a.reps)
    barrier()
    mine = {"rank": rank, "alone_h2d_GBps": alone[rank][0], "alone_ms": alone[rank][2],
            "together_h2d_GBps": together[0], "together_ms": together[2], "bound_cpus": bound,
            "cpus_allowed": len(os.sched_getaffinity(0))}
    if world > 1:
        allr = [None] * world
        dist.all_gather_object(allr, mine)
    else:
        allr = [mine]
    if rank == 0:
        print(json.dumps({"n_ranks": world, "block_MB": n_in * 4 / 1e6, "bind": a.bind, "ranks": allr,
                          "sum_together_GBps": sum(r["together_h2d_GBps"] for r in allr),
                          "host_cpus": os.cpu_count(), "topo": t}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
