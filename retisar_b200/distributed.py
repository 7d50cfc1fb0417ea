"""Multi-GPU plumbing: streams shard across ranks, nothing is exchanged on the render path.

One process per GPU (``torch.distributed``); every rank renders its own contiguous range of
streams with replicated tables.  The only collectives are a max-reduction of
timings and an optional gather of result blocks to rank 0, both off the hot path.
"""


def stream_shard(n_streams_total, world_size, rank):
    """Contiguous range [start, stop) of global stream ids rendered by `rank`."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    base, extra = divmod(n_streams_total, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def max_over_ranks(value, device=None):
    """Maximum of a python float over all ranks (identity without an initialised group)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_streams(block, dst=0):
    """Gather per-rank result blocks [S_local, E, B] to `dst`; returns [S_total, E, B] there,
    None elsewhere.  Ranks may hold different numbers of streams."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return block
    world, rank = dist.get_world_size(), dist.get_rank()
    counts = [torch.zeros(1, dtype=torch.int64, device=block.device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([block.shape[0]], dtype=torch.int64, device=block.device))
    n_max = int(max(c.item() for c in counts))
    padded = torch.zeros((n_max,) + tuple(block.shape[1:]), dtype=block.dtype, device=block.device)
    padded[: block.shape[0]] = block
    parts = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, parts, dst=dst)
    if rank != dst:
        return None
    return torch.cat([p[: int(c.item())] for p, c in zip(parts, counts)], dim=0)


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index):
    """Restrict this process to the CPUs of the NUMA node the GPU is attached to, so that the
    page-locked buffers it allocates afterwards (first touch) and the thread that feeds them sit
    next to the GPU's PCIe root port.  With one process per GPU this keeps every rank's host-side
    block traffic on its own socket instead of all ranks sharing node 0.

    Returns a dict describing what was done (``node`` is None when the platform exposes no
    topology, e.g. inside a VM: nothing is changed then).
    """
    import os

    info = {"node": None, "cpus": None, "bdf": None}
    try:
        import torch

        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    except Exception:  # older torch: ask the driver
        try:
            import subprocess

            bdf = subprocess.run(
                ["nvidia-smi", f"--id={device_index}", "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                stdout=subprocess.PIPE, text=True, timeout=10).stdout.strip().lower()
            if len(bdf.split(":")[0]) == 8:  # nvidia-smi prints an 8-digit domain
                bdf = bdf[4:]
        except Exception:
            return info
    info["bdf"] = bdf
    try:
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read())
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info.update(node=node, cpus=len(cpus))
    except (OSError, ValueError):
        pass
    return info
