#!/usr/bin/env python
"""Benchmark of the per-block binaural rendering chain (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c2|c5|c3|c4|c1] [--streams S] [--no-extras]

A "step" is one block (256 samples at 48 kHz for c2) of S batched, independent render streams
with their own head yaw.  One JSON line is printed by rank 0:

  value      real-time streams sustained = stream-blocks per second / (fs / B), whole job,
             inputs resident in HBM, exactly K steps back to back (device time, max over ranks)
  e2e        same metric through the public host-buffer API (`AdjustableShConvolver.submit_block`
             / `.result()`, pinned HOST buffers): H2D copy of every block, render, D2H copy of the
             ear signals inside the timed region; `e2e.sync` = plain `filter_block` per block
  roofline   dominant kernel (the render kernel), timed live with CUDA events on its stream
  sweep      S in {1024, 4096, 10240}: ms per step and p99 block latency (>= 500 latency steps)
  held       largest S whose p99 block latency stays below 1 ms / below the block period
             (the reference's own criterion "no dropouts", `_run_benchmark.py:311-321`)
  fdl        roofline of the FDL kernel `ols_fdl_kernel` (north-star target: >= 60 % of HBM)
  scaling_c5 BASELINE config 5 (order 8, 8192 streams per GPU) measured in the same run
  c3_partitioned / c4_chain / c1_single_stream
             BASELINE configs 3 (16-partition HRIR, 4096 streams), 4 (production chain ARIR ->
             renderer -> HPCF, 1024 streams) and 1 (one stream) measured in the same run
  cpu_baseline / --impl reference
             the UNMODIFIED reference's `AdjustableShConvolver.filter_block` (staged into
             oracle/_ref by build(); kind "reference") on the host cores, one process per core;
             falls back to the numpy oracle (kind "port") when no staged copy exists
"""
import argparse
import json
import os

# the reference pins BLAS / OpenMP to one thread per process (`ReTiSAR/config.py:277-280`,
# `tools.py:416-460`); must happen before numpy is imported
for _var in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ.setdefault(_var, "1")
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "real-time binaural streams @48kHz/256-blk per GPU; p99 block latency"
UNIT = "realtime_streams"
HRIR_DIRS = 2702  # the reference's default HRIR set (`config.py:48`), both arms


def workload_config(workload, streams):
    """`config` of the JSON line -- identical for both arms (the driver compares them)."""
    from retisar_b200 import workloads as wl

    c = wl.CONFIGS[workload]
    return {"workload": f"{workload}: {c['desc']}", "streams_per_gpu": streams,
            "block_length": c["block"], "fs": wl.FS, "sh_order": c["order"],
            "hrir_directions": HRIR_DIRS, "crossfade": True}


# ------------------------------------------------------------------------------------------------
# CPU arm: the unmodified reference (or its numpy port), one process per host core
# ------------------------------------------------------------------------------------------------
def _reference_kind(workload):
    """"reference" when the unmodified package is importable and supports the workload."""
    from retisar_b200 import workloads as wl

    try:
        from oracle import ref_harness as rh
    except ImportError:
        return "port"
    c = wl.CONFIGS[workload]
    if not rh.reference_available() or c["hrir_taps"] > c["block"]:  # the reference refuses P > 1
        return "port"
    return "reference"


def _cpu_worker(args):
    """Render `n_blocks` blocks of one stream on one core; returns seconds per block."""
    workload, n_blocks, warm, seed, kind = args
    from retisar_b200 import synthetic as syn
    from retisar_b200 import workloads as wl

    c = wl.CONFIGS[workload]
    B, order = c["block"], c["order"]
    az, co, w = wl._grid(c["grid"])
    haz, hco = syn.random_sphere_grid(HRIR_DIRS, 7)
    hrir_td = syn.decaying_noise_irs((HRIR_DIRS, 2, c["hrir_taps"]), 8)
    x = syn.noise_blocks(8, len(az), B, seed)
    yaw = syn.yaw_walk(warm + n_blocks, seed + 1)
    if kind == "reference":
        # the reference's own classes, imported unmodified (oracle/ref_harness.py)
        from oracle import ref_harness as rh

        ref = rh.import_reference()
        arr = rh.make_miro_filter_set(np.zeros((1, len(az), B), np.float32), az, co, w, order,
                                      is_hrir=False, block_length=B)
        hrir = rh.make_miro_filter_set(hrir_td, haz, hco, None, order, is_hrir=True, block_length=B)
        hrir.calculate_filter_blocks_fd(B)
        tracker = rh.make_tracker()
        conv = rh.make_sh_convolver(hrir, arr, B, tracker)
        azim = ref.HeadTracker.DataIndex.AZIM

        def block(i, y):
            tracker[azim] = float(y)
            return conv.filter_block(x[i % 8].copy())  # the JACK client hands over a fresh copy
    else:
        from oracle import sh_chain as oc

        tb = wl.sh_tables(workload, hrir_dirs=HRIR_DIRS)
        cfg = tb["sh_config"]
        tb["hrir_set"].calculate_filter_blocks_fd(B)
        tb["hrir_set"].calculate_filter_blocks_nm()
        hnm = tb["hrir_set"].get_filter_blocks_nm()
        cls = oc.ShOracle if hnm.shape[0] == 1 else oc.ShPartitionedOracle
        o = cls(cfg.sh_m, cfg.sh_bases_weighted, hnm, B, order)

        def block(i, y):
            return o.filter_block(x[i % 8], y)
    for i in range(warm):
        block(i, yaw[i])
    t0 = time.perf_counter()
    for i in range(n_blocks):
        block(i, yaw[warm + i])
    return (time.perf_counter() - t0) / n_blocks


def cpu_arm(workload, cores, n_blocks, warm=50):
    """Streams the host sustains in real time with `cores` independent single-stream processes."""
    import multiprocessing as mp

    from retisar_b200 import workloads as wl

    kind = _reference_kind(workload)
    block_period = wl.CONFIGS[workload]["block"] / wl.FS
    jobs = [(workload, n_blocks, warm, i, kind) for i in range(cores)]
    if cores == 1:
        per_block = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            per_block = pool.map(_cpu_worker, jobs)
    streams = sum(block_period / t for t in per_block)
    sec = float(np.mean(per_block))
    what = ("the UNMODIFIED reference AdjustableShConvolver.filter_block (oracle/_ref, np.fft branch)"
            if kind == "reference" else "the numpy oracle (port of the reference)")
    sample = (f"{cores} processes x 1 stream x {n_blocks} blocks of {what}, BLAS threads = 1, "
              f"{HRIR_DIRS} HRIR directions, {sec * 1e6:.0f} us/block/stream")
    return streams, sec, kind, sample


def run_reference_arm(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    # One step = a bounded sample: every core renders `per_step` consecutive blocks of its own stream.
    # W warm-up steps run untimed, K steps are timed; the sample shrinks when K is large so that the
    # whole run stays below about a minute of wall time (the set-up of the 2702-direction HRIR set is
    # a few seconds on top).
    per_step = max(20, min(1000, 200000 // max(args.steps, 1)))
    n_blocks, warm = args.steps * per_step, max(args.warmup, 1) * per_step
    t0 = time.perf_counter()
    streams, sec, kind, sample = cpu_arm(args.workload, cores, n_blocks, warm=warm)
    wall = time.perf_counter() - t0
    line = {
        "impl": "reference", "metric": METRIC, "value": streams, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, args.streams),
        "cpu_baseline": {"value": streams, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": streams, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "step_definition": f"one step = a bounded sample of {per_step} consecutive blocks of one stream on "
                           f"every core ({args.warmup} untimed + {args.steps} timed steps); ms_per_step = mean "
                           f"time of ONE block of one stream, as in the GPU arm (there: one block of the batch)",
        "wall_s": wall,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([v.strip() for v in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); smax.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class Rig:
    """One renderer with S streams and its HBM-resident input ring (larger than L2)."""

    def __init__(self, tb, S, dev, seed, min_ring_bytes=160e6):
        import torch

        from retisar_b200 import workloads as wl

        self.tb, self.S, self.dev = tb, S, dev
        self.B, self.C, self.E = tb["B"], tb["C"], 2
        self.conv = wl.make_renderer(tb, S, device=dev.index)
        self.plan = self.conv._plan
        self.in_bytes = S * self.C * self.B * 4
        self.ring = max(2, int(np.ceil(min_ring_bytes / self.in_bytes)) + 1)
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed)
        self.x_ring = torch.empty((self.ring, S, self.C, self.B), device=dev, dtype=torch.float32)
        for r in range(self.ring):  # block by block: no second copy of a multi-GB ring
            self.x_ring[r].normal_(0.0, 0.1, generator=gen)
        yaw0 = torch.rand((S,), generator=gen, device=dev) * 360.0
        self.yaw_ring = torch.stack([yaw0 + 1.5 * i for i in range(self.ring)]).contiguous()
        self.out = torch.empty((S, self.E, self.B), device=dev, dtype=torch.float32)
        # an explicit stream (made torch's current one, so that the L2 flushes and every other torch
        # operation of the bench stay ordered with the launches): the library replays a repeated run of
        # blocks as a CUDA graph only on an explicit stream
        if not getattr(Rig, "_stream", None):
            Rig._stream = torch.cuda.Stream(dev)
            torch.cuda.set_stream(Rig._stream)
        self.stream = Rig._stream
        self.n_done = 0

    def run(self, count):
        """`count` consecutive blocks enqueued by ONE library call (no Python between the launches)."""
        self.plan.process_blocks_device(self.x_ring, self.yaw_ring, self.out, self.n_done, count, self.stream)
        self.n_done += count

    def run_from_ring_start(self, count):
        """The same run every time (ring phase 0): from its third occurrence on the library replays it
        as one CUDA graph launch (`rtb_sh_process_blocks`)."""
        self.plan.process_blocks_device(self.x_ring, self.yaw_ring, self.out, 0, count, self.stream)

    def timed(self, count, primed=False):
        import torch

        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        if primed:
            self.run_from_ring_start(count)
        else:
            self.run(count)
        e1.record(self.stream)
        return e0, e1

    def latencies(self, n, flush=None):
        """Per-block device time of n blocks, one at a time; L2 is flushed before every block
        unless a single input block already exceeds it."""
        import torch

        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
               for _ in range(n)]
        for a, b in evs:
            if flush is not None:
                flush.add_(1.0)
            a.record(self.stream)
            self.run(1)
            b.record(self.stream)
        torch.cuda.synchronize(self.dev)
        return np.array([a.elapsed_time(b) for a, b in evs])

    def close(self):
        self.conv._plan.close()
        del self.x_ring, self.yaw_ring, self.out


def _percentiles(lat):
    return {k: float(np.percentile(lat, q)) for k, q in (("p50", 50), ("p99", 99), ("max", 100))}


def measure_fdl(dev, peaks, S, taps, B, n_ch=2, steps=100):
    """Roofline of `ols_fdl_kernel` on the reference benchmark's own shape (2-in / 2-out FIR)."""
    import torch

    from retisar_b200 import _capi
    from retisar_b200.plans import OlsPlan

    P, F = -(-taps // B), B + 1
    Fp = (F + 3) // 4 * 4
    rng = np.random.default_rng(0)
    irs = (rng.standard_normal((n_ch, P * B)) * 0.05 * np.exp(-np.arange(P * B) / (P * B / 6))).astype(np.float32)
    H = np.stack([np.fft.rfft(irs[:, p * B:(p + 1) * B], 2 * B) for p in range(P)]).astype(np.complex64)
    plan = OlsPlan(S, B, n_ch, n_ch, P, device=dev.index)
    plan.set_filter(H)
    ring = max(2, int(np.ceil(160e6 / (S * n_ch * B * 4))) + 1)
    x = torch.randn((ring, S, n_ch, B), device=dev) * 0.1
    out = torch.empty((S, n_ch, B), device=dev)
    st = torch.cuda.current_stream(dev)
    for i in range(P + 4):
        plan.process_device(x[i % ring], out, st)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for i in range(steps):
        plan.process_device(x[i % ring], out, st)
    e1.record(st)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    plan.set_profiling(True)
    ks = []
    for i in range(30):
        plan.process_device(x[i % ring], out, st)
        ks.append(plan.query(_capi.RTB_Q_RENDER_NS))
    plan.close()
    us = float(np.mean(ks)) / 1e3
    # input-side delay line: read P stored spectra, write one; previous + current input block in,
    # state and output out
    bytes_per = 8 * Fp * n_ch * P + 4 * B * (3 * n_ch + n_ch)
    ach = S * bytes_per / (us * 1e-6) / 1e9
    return {"kernel": "ols_fdl_kernel", "shape": f"{n_ch}-in/{n_ch}-out, {taps} taps, B = {B}, P = {P}",
            "streams": S, "bound": "hbm", "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s",
            "frac": ach / peaks["hbm"], "bytes_per_launch": S * bytes_per, "us_per_launch": us,
            "ms_per_step": ms, "realtime_convolvers": S / (ms * 1e-3) / (48000 / B),
            "traffic": peaks["traffic"].get(f"fdl:{taps}:{B}:{S}", {}).get("bytes")}


def run_gpu_arm(args):
    import torch

    from retisar_b200 import _capi
    from retisar_b200 import workloads as wl
    from retisar_b200.distributed import bind_to_gpu_numa_node, gather_streams, max_over_ranks as _max

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU path)")
    orig_affinity = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # host side of the e2e leg: this rank's feeder thread and its page-locked buffers next to its GPU
    numa = bind_to_gpu_numa_node(local_rank) if not args.no_numa_bind else {"node": None}
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        return _max(v, device=dev)

    peaks_file = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks_file = json.load(f)
    except OSError:
        pass
    traffic_tab = {}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic_tab = json.load(f)
    except (OSError, ValueError):
        pass
    peaks = {"hbm": float(peaks_file.get("hbm_gbs", 6650.0)), "traffic": traffic_tab,
             "source": "measured (MEASURED_PEAKS.json)" if peaks_file else "fallback (B200_PROFILING.md)"}

    S, K, W = args.streams, args.steps, max(args.warmup, 3)
    tb = wl.sh_tables(args.workload, hrir_dirs=HRIR_DIRS)
    B, C, E = tb["B"], tb["C"], 2
    block_rate = wl.FS / B  # blocks per second of one real-time stream
    flush = torch.empty(int(256e6) // 4, device=dev, dtype=torch.float32)
    rig = Rig(tb, S, dev, 1234 + rank)
    plan, conv = rig.plan, rig.conv
    rig_stream = rig.stream

    # ---- device-resident throughput: W warm-up, then exactly K steps between barriers
    rig.run(W)
    # the run of K steps that is timed below, twice, untimed: the library records a run that comes back
    # unchanged as a CUDA graph (second occurrence) and replays it from then on -- the timed K steps
    # are one graph launch, no host between the kernels (four occurrences for an odd K: the ping-pong
    # state of the plan alternates between two keys)
    for _ in range(2 if K % 2 == 0 else 4):
        rig.run_from_ring_start(K)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = plan.query(_capi.RTB_Q_LAUNCHES)
    replays0 = plan.query(_capi.RTB_Q_GRAPH_REPLAYS)
    e0, e1 = rig.timed(K, primed=True)
    barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    launches = plan.query(_capi.RTB_Q_LAUNCHES) - launches0
    graph_replays = plan.query(_capi.RTB_Q_GRAPH_REPLAYS) - replays0
    value = world * S * K / (dev_ms * 1e-3) / block_rate

    # ---- block latency: per-step events, L2 flushed before every step
    n_lat = max(500, min(K, 2000))
    lat = rig.latencies(n_lat, flush)
    pct = _percentiles(lat)
    lat_p99 = max_over_ranks(pct["p99"])

    # ---- per-kernel device time (CUDA events inside the library, on the launching stream); blocks
    # one at a time from the same > L2 input ring as the timed loop, no flush in between
    plan.set_profiling(True)
    k_an, k_re = [], []
    for i in range(100):
        rig.run(1)
        k_an.append(plan.query(_capi.RTB_Q_ANALYSIS_NS))
        k_re.append(plan.query(_capi.RTB_Q_RENDER_NS))
    plan.set_profiling(False)
    an_us, re_us = float(np.mean(k_an)) / 1e3, float(np.mean(k_re)) / 1e3

    # ---- end to end through the public API with pinned host buffers
    K_e2e = max(K, 50)
    hx = torch.randn((4, S, C, B), dtype=torch.float32).mul_(0.1).pin_memory()
    hyaw = (torch.rand((4, S)) * 360.0).pin_memory()
    hx_np, hyaw_np = hx.numpy(), hyaw.numpy()
    def e2e_sync(n):  # the reference-shaped synchronous call, one block at a time
        acc = 0.0
        for i in range(n):
            y = conv.filter_block(hx_np[i % 4], hyaw_np[i % 4])
            acc += float(y[0, 0, 0])
        return acc

    def e2e_pipelined(n):  # feeder shape: block t+1 is handed over before block t is collected
        acc = 0.0
        pending = conv.submit_block(hx_np[0], hyaw_np[0])
        for i in range(1, n + 1):
            nxt = conv.submit_block(hx_np[i % 4], hyaw_np[i % 4]) if i < n else None
            y = pending.result()
            acc += float(y[0, 0, 0])
            pending = nxt
        return acc

    def timed_host(fn, n):
        barrier()
        t0 = time.perf_counter()
        fn(n)
        barrier()
        return max_over_ranks(time.perf_counter() - t0)

    # warm-up of both call shapes (staging slots, pinned result pool, link out of its idle state), then
    # E2E_REPS repetitions of K_e2e steps each; the line carries the MEDIAN repetition and lists all of
    # them (one repetition is 30 - 40 ms of wall time: a single one is at the mercy of one host hiccup;
    # fresh boxes were seen to deliver 0.64 - 0.94 ms per step in the first repetition of a process)
    e2e_sync(8)
    e2e_pipelined(8)
    E2E_REPS = 5
    sync_reps = [timed_host(e2e_sync, K_e2e) for _ in range(E2E_REPS)]
    pipe_reps = [timed_host(e2e_pipelined, K_e2e) for _ in range(E2E_REPS)]
    e2e_sync_s, e2e_s = float(np.median(sync_reps)), float(np.median(pipe_reps))
    e2e_value = world * S * K_e2e / e2e_s / block_rate
    clocks = sampler.stop() if sampler else None

    # ---- gather one output block from every rank (the only collective; off the timed path)
    gathered = gather_streams(rig.out, dst=0)
    if rank == 0 and dist is not None:
        assert gathered.shape[0] == world * S and bool(torch.isfinite(gathered).all())

    J = plan.query(_capi.RTB_Q_ROWS)
    render_name = _capi.RENDER_KERNEL_NAMES.get(plan.query(_capi.RTB_Q_RENDER_KERNEL), "sh_render_kernel")
    analysis_name = "sh_analysis_tc_kernel" if plan.query(_capi.RTB_Q_TENSOR_CORE) else "sh_analysis_kernel"
    in_ring_desc = (f"inputs cycle through a {rig.ring * rig.in_bytes / 1e6:.0f} MB ring of {rig.ring} blocks "
                    f"(> 126 MB L2); latency steps flush L2 with a 256 MB write")
    rig.close()

    # ---- extras measured in the same run --------------------------------------------------
    extras = {}
    if not args.no_extras:
        def quick(workload, S_x, n_steps, n_lat_x, seed):
            tbx = tb if workload == args.workload else wl.sh_tables(workload, hrir_dirs=HRIR_DIRS)
            r = Rig(tbx, S_x, dev, seed)
            r.run(max(5, W))
            torch.cuda.synchronize(dev)
            a, b = r.timed(n_steps)
            torch.cuda.synchronize(dev)
            ms = a.elapsed_time(b) / n_steps
            fl = flush if r.in_bytes < 256e6 else None
            p = _percentiles(r.latencies(n_lat_x, fl)) if n_lat_x else None
            period = 1e3 * tbx["B"] / wl.FS
            r.close()
            return ms, p, period

        # BASELINE config 5 beside the headline, at every N (weak scaling: 8192 streams per GPU)
        barrier()
        ms5, p5, period5 = quick("c5", 8192, 30, 100, 77 + rank)
        ms5 = max_over_ranks(ms5)
        extras["scaling_c5"] = {
            "workload": "c5: " + wl.CONFIGS["c5"]["desc"], "streams_per_gpu": 8192,
            "value": world * 8192 / (ms5 * 1e-3) / (wl.FS / 256), "unit": UNIT, "ms_per_step": ms5,
            "p99_block_latency_ms": max_over_ranks(p5["p99"]), "steps": 30, "latency_steps": 100}
        if world == 1 and args.workload == "c2":
            sweep = {}
            for S_x in (1024, 4096, 10240):
                ms, p, period = quick("c2", S_x, 200, 500, 5 + S_x)
                sweep[str(S_x)] = {"ms_per_step": ms, "p99_ms": p["p99"], "p50_ms": p["p50"],
                                   "max_ms": p["max"], "latency_steps": 500,
                                   "streams_sustained": S_x / (ms * 1e-3) / block_rate}
            extras["sweep"] = sweep

            def holds(S_x, limit_ms, n):
                _, p, _ = quick("c2", S_x, 20, n, 11 + S_x)
                return p["p99"] < limit_ms, p

            def search(limit_ms, start, step, cap):
                """Largest S (multiple of `step`) whose p99 over 300 blocks stays below the limit,
                then confirmed over 2000 consecutive blocks (stepping down while it fails)."""
                lo, hi = 0, None
                s = start
                while hi is None and s <= cap:
                    ok, _ = holds(s, limit_ms, 300)
                    if ok:
                        lo, s = s, int(s * 1.5) // step * step
                    else:
                        hi = s
                if hi is None:
                    hi = cap + step
                while hi - lo > step:
                    mid = (lo + hi) // 2 // step * step
                    if mid <= lo:
                        break
                    ok, _ = holds(mid, limit_ms, 300)
                    lo, hi = (mid, hi) if ok else (lo, mid)
                while lo > 0:
                    ok, p = holds(lo, limit_ms, 2000)
                    if ok:
                        return {"streams": lo, "p99_ms": p["p99"], "p50_ms": p["p50"], "max_ms": p["max"],
                                "limit_ms": limit_ms, "confirmed_blocks": 2000}
                    lo -= step
                return {"streams": 0, "limit_ms": limit_ms}

            extras["held"] = {
                "criterion": "largest S with p99(block device time) below the limit over 2000 consecutive "
                             "blocks, one block per launch pair, L2 flushed (or input block > L2) "
                             "between blocks; search in steps of 512 (1 ms) / 4096 (block period)",
                "p99_below_1ms": search(1.0, 10240, 512, 200000),
                "p99_below_block_period": search(1e3 / block_rate, 98304, 4096, 400000)}
            # BASELINE config 1: ONE em32 / order-4 stream, the case the reference runs on a CPU core
            ms1, p1, _ = quick("c1", 1, 200, 500, 3)
            one = wl.make_renderer(tb, 1, device=dev.index)
            xb = (np.random.default_rng(0).standard_normal((C, B)) * 0.1).astype(np.float32)
            for i in range(20):
                one.filter_block(xb, [10.0 + i])
            host_lat = []
            for i in range(300):
                t1 = time.perf_counter()
                one.filter_block(xb, [30.0 + i])
                host_lat.append((time.perf_counter() - t1) * 1e3)
            one._plan.close()
            extras["c1_single_stream"] = {
                "device_ms_per_block": ms1, "device_p99_ms": p1["p99"],
                "filter_block_host_p50_ms": float(np.percentile(host_lat, 50)),
                "filter_block_host_p99_ms": float(np.percentile(host_lat, 99)),
                "note": "one stream, block latency; filter_block_host = wall time of the synchronous "
                        "reference-shaped call with a [32,256] numpy block (H2D + 2 kernels + D2H)"}
            # BASELINE config 3: order 8, 4096-tap HRIR as 16 partitions (extension a-X), 4 096 streams
            from retisar_b200 import config as rcfg

            rcfg.IS_SH_TRANSFORM_ON_DEVICE = True  # 2 702 directions x 16 partitions: set-up on the tensor cores
            try:
                ms3, p3, _ = quick("c3", 4096, 30, 100, 9)
            finally:
                rcfg.IS_SH_TRANSFORM_ON_DEVICE = False
            extras["c3_partitioned"] = {
                "workload": "c3: " + wl.CONFIGS["c3"]["desc"], "streams_per_gpu": 4096, "ms_per_step": ms3,
                "value": 4096 / (ms3 * 1e-3) / block_rate, "unit": UNIT, "p99_block_latency_ms": p3["p99"],
                "steps": 30, "latency_steps": 100,
                "note": "partition MAC on tcgen05 (sh_partition_tc_kernel); SH transform of the HRIR set on the device"}
            # BASELINE config 4 as the reference chains it in production: mono source -> 434-channel ARIR
            # (4096 taps) -> renderer (order 12, 64-sample blocks) -> HPCF (1024 taps), folded into one renderer
            tb4 = wl.sh_tables("c4", hrir_dirs=HRIR_DIRS)
            from retisar_b200 import FilterSet
            from retisar_b200 import synthetic as syn

            rend4 = wl.make_renderer(tb4, 1024, device=dev.index)
            rend4.set_pre_renderer(syn.decaying_noise_irs((tb4["C"], 4096), 40))
            hp4 = FilterSet.create_instance_by_type(syn.decaying_noise_irs((2, 1024), 41, scale=0.3), "FIR_MULTICHANNEL")
            hp4.load(block_length=tb4["B"], is_single_precision=True, is_prevent_logging=True)
            rend4.set_hpcf(hp4)
            x4 = torch.randn((16, 1024, 1, tb4["B"]), device=dev) * 0.5
            y4 = torch.rand((16, 1024), device=dev) * 360.0
            for i in range(80):
                rend4.filter_block(x4[i % 16], y4[i % 16])
            torch.cuda.synchronize(dev)
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record(rig_stream)
            for i in range(100):
                rend4.filter_block(x4[i % 16], y4[i % 16])
            eb.record(rig_stream)
            torch.cuda.synchronize(dev)
            ms4 = ea.elapsed_time(eb) / 100
            extras["c4_chain"] = {
                "workload": "c4 production chain: mono source -> 434-ch ARIR pre-renderer (4096 taps, 64 partitions) -> "
                            "SH renderer (order 12, 64-sample blocks) -> HPCF (1024 taps, 16 partitions)",
                "streams_per_gpu": 1024, "ms_per_step": ms4, "block_period_ms": 1e3 * tb4["B"] / wl.FS,
                "value": 1024 / (ms4 * 1e-3) / (wl.FS / tb4["B"]), "unit": "realtime_streams @48kHz/64-blk",
                "steps": 100,
                "note": "ARIR folded into the analysis domain (set_pre_renderer), HPCF fused into the render kernel "
                        "(set_hpcf); the three separate convolvers take 0.449 ms (profiles/r02_c4_chain_*)"}
            del rend4, x4, y4
            extras["fdl"] = {
                "target": ">= 60 % of the measured HBM peak on the FDL kernel (north star)",
                "taps4096_b256": measure_fdl(dev, peaks, 16384, 4096, 256),
                "hpcf_1024taps_b64": measure_fdl(dev, peaks, 16384, 1024, 64)}

    if rank == 0:
        render_bytes = S * (4 * B * (2 * J + E) + 12)  # z rows of 2 blocks in, ear blocks out, yaw
        P = -(-wl.CONFIGS[args.workload]["hrir_taps"] // B)
        if P > 1:  # a-X: plus both output-side queues [P][E][F] c64, read and written once per block
            render_bytes += S * 2 * 2 * 8 * P * E * (B + 1)
        analysis_bytes = S * 4 * B * (C + J)
        chain_bytes = S * wl.chain_bytes_per_stream_block(C, B, E)
        chain_flops = S * wl.chain_flops_per_stream_block(C, tb["K"], B, E)
        achieved = render_bytes / (re_us * 1e-6) / 1e9
        traffic = traffic_tab.get(f"{args.workload}:{S}", {}).get("bytes")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, S),
            "timing": {"input_ring_blocks": rig.ring, "l2_policy": in_ring_desc, "channels": C,
                       "launch": "the K timed steps are enqueued by one rtb_sh_process_blocks call",
                       "numa": numa},
            "p99_block_latency_ms": lat_p99, "p50_block_latency_ms": pct["p50"],
            "max_block_latency_ms": pct["max"], "block_period_ms": 1e3 / block_rate,
            "latency_steps": n_lat,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": S * C * B * 4 + S * 4,
                    "d2h_bytes_per_step": S * E * B * 4, "steps": K_e2e,
                    "ms_per_step": e2e_s / K_e2e * 1e3,
                    "statistic": f"median of {E2E_REPS} repetitions of {K_e2e} steps (each between barriers, "
                                 f"max over ranks); all repetitions in reps_ms_per_step",
                    "reps_ms_per_step": [t / K_e2e * 1e3 for t in pipe_reps],
                    "api": "AdjustableShConvolver.submit_block(numpy pinned [S,C,B], yaw[S]) / ticket.result(): "
                           "block t+1 submitted before block t is collected (feeder shape)",
                    "sync": {"value": world * S * K_e2e / e2e_sync_s / block_rate,
                             "ms_per_step": e2e_sync_s / K_e2e * 1e3,
                             "reps_ms_per_step": [t / K_e2e * 1e3 for t in sync_reps],
                             "api": "AdjustableShConvolver.filter_block(numpy pinned [S,C,B], yaw[S])"}},
            "gpu_launches": int(launches),
            "launch_mode": ("the K timed steps were ONE CUDA graph launch (run recorded by the library during two "
                            "untimed occurrences of the same run after the W warm-up steps)" if graph_replays == 1
                            else "the kernels of the K timed steps were launched one by one from one C call"),
            "graph_replays_in_timed_region": int(graph_replays),
            "roofline": {"bound": "hbm", "kernel": render_name, "achieved": achieved,
                         "peak": peaks["hbm"], "unit": "GB/s", "frac": achieved / peaks["hbm"],
                         "traffic": traffic, "peak_source": peaks["source"],
                         "bytes_per_launch": render_bytes, "us_per_launch": re_us,
                         "note": "HBM figure as the contract asks; ncu (profiles/) shows this kernel bound by "
                                 "the L1/shared-memory data pipe and the FP32 pipe, not by HBM; the HBM-bound "
                                 "kernels are the analysis GEMM ('kernels') and the FDL kernel ('fdl').  "
                                 "us_per_launch: library events around single blocks read from the same > L2 "
                                 "input ring as the timed loop (no extra flush)"},
            "kernels": {analysis_name: {"us": an_us, "bytes": analysis_bytes,
                                        "GBps": analysis_bytes / (an_us * 1e-6) / 1e9,
                                        "frac_hbm": analysis_bytes / (an_us * 1e-6) / 1e9 / peaks["hbm"]},
                        render_name: {"us": re_us, "bytes": render_bytes, "GBps": achieved}},
            "chain": {"bytes_per_step": chain_bytes, "flops_per_step": chain_flops,
                      "GBps": chain_bytes / (dev_ms / K * 1e-3) / 1e9,
                      "TFLOPs_reference_formulation": chain_flops / (dev_ms / K * 1e-3) / 1e12},
        }
        line.update(extras)
        if world == 1 and not args.no_cpu_baseline:
            if orig_affinity is not None:  # every core the process started with (undo the NUMA binding)
                os.sched_setaffinity(0, orig_affinity)
            cores = len(orig_affinity) if orig_affinity is not None else (os.cpu_count() or 1)
            # about 1 s of wall time per core: 16 cores -> ~18 s of CPU work (bounded sample)
            cb_streams, cb_sec, kind, sample = cpu_arm(args.workload, cores, 8000)
            line["cpu_baseline"] = {"value": cb_streams, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": sample}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c5", "c4", "b4096"])
    ap.add_argument("--streams", type=int, default=1024, help="streams per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip sweep / held / fdl / scaling_c5 (only the headline line)")
    ap.add_argument("--no-numa-bind", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
