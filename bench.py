#!/usr/bin/env python
"""Benchmark of the propagation hot path (BASELINE.json metric: SPN propagation Gpix.iter/s & % HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A *step* is one training pass of the hot path over one batch of synthetic DFC30-shaped tiles:
PostProcessor forward + backward (gradients for the affinities, the offsets and the 3x3 weight / bias; the DEM is
detached exactly as models/JSPSR.py:372 does) on configs/jspsr_r8_img.yml's layer (3x3, residual, 128x128 tiles).
The per-GPU batch is 4096 tiles (67 Mpix, 7.8 GB of inputs - far larger than the 126 MB L2, so nothing is
cache-resident between steps); the YAML batch of 70 tiles is launch-latency bound and is
reported separately in `config_batch`.  Multi-GPU: one process per GPU, the batch is sharded (weak scaling), the only
cross-rank state of the path - grad_w[9] and grad_b[1] - is all-reduced over NCCL inside the step.

`--impl reference` times the reference's CPU implementation of the same step on the host cores: the restated call
sites of spn.py:99-118 on torchvision's own CPU operator (oracle/ref_port.py) - literally what the reference executes
on CPU - with all host threads, on a bounded sample of the same workload.  The fused OpenMP C restatement
(oracle/spn_oracle.c), which is a different and much faster CPU program than the reference's, is reported beside it
(`c_restatement`) and is the fallback when torchvision is not importable.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "SPN propagation Gpix·iter/s & % HBM roofline"
UNIT = "Gpix·iter/s"
TILE = 128
FWD_BYTES = {"f32": 116, "bf16": 58}    # per pixel per application, SURVEY.md section 8d
BWD_BYTES = {"f32": 224, "bf16": 112}   # grad_init not required (detached DEM)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="tiles per GPU per step")
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-tiles", type=int, default=32, help="tiles in the bounded CPU sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    return ap.parse_args()


def workload_config(args, n):
    return {
        "workload": f"configs/jspsr_r8_img.yml PostProcessor(3x3, residual) training step fwd+bwd, "
                    f"{args.batch} tiles of {TILE}x{TILE} per GPU, T=1, DEM detached",
        "tiles_per_gpu": args.batch, "tile": [TILE, TILE], "global_tiles": args.batch * n,
        "parallelism": f"dp{n} (batch-sharded tiles, NCCL all-reduce of grad_w/grad_b)",
        "l2": "inputs (7.8 GB/GPU) exceed the 126 MB L2; no flush needed",
        "offsets": "N(0,1.5^2) clipped to +-8, centre pair zero (SURVEY.md section 8d)",
    }


def peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic(kernel):
    """dram bytes per launch from the committed ncu --set full capture of this workload, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------------------
# clocks sampled DURING the timed region
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
                 "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------------
# synthetic inputs
# ---------------------------------------------------------------------------------------------------------
def make_inputs(torch, B, device, dtype, seed):
    g = torch.Generator(device=device).manual_seed(seed)
    init = torch.rand(B, 1, TILE, TILE, device=device, generator=g)
    weight = torch.sigmoid(1.5 * torch.randn(B, 9, TILE, TILE, device=device, generator=g))
    offset = (1.5 * torch.randn(B, 18, TILE, TILE, device=device, generator=g)).clamp_(-8, 8)
    offset[:, 8:10] = 0
    gout = torch.randn(B, 1, TILE, TILE, device=device, generator=g)
    w = torch.ones(1, 1, 3, 3, device=device) + (torch.rand(1, 1, 3, 3, device=device, generator=g) - 0.5) * 0.2
    b = torch.full((1,), 0.1, device=device)
    return [t.to(dtype) for t in (init, weight, offset, gout)] + [w, b]


# ---------------------------------------------------------------------------------------------------------
# CPU baseline / reference arm
# ---------------------------------------------------------------------------------------------------------
def cpu_step_fns(tiles):
    """Two CPU implementations of the same step; returns {name: (callable, cores)}."""
    import numpy as np
    import torch
    rng = np.random.default_rng(1234)
    init = rng.random((tiles, 1, TILE, TILE), dtype=np.float32)
    weight = (1 / (1 + np.exp(-1.5 * rng.normal(size=(tiles, 9, TILE, TILE))))).astype(np.float32)
    offset = np.clip(1.5 * rng.normal(size=(tiles, 18, TILE, TILE)), -8, 8).astype(np.float32)
    offset[:, 8:10] = 0
    gout = rng.normal(size=(tiles, 1, TILE, TILE)).astype(np.float32)
    w9 = (1 + 0.2 * (rng.random(9) - 0.5)).astype(np.float32)
    b1 = np.array([0.1], np.float32)
    fns = {}
    try:
        from oracle import ref_port
        torch.set_num_threads(os.cpu_count() or 1)
        t = [torch.from_numpy(a) for a in (init, weight, offset)]
        tw, tb, tg = torch.from_numpy(w9.reshape(1, 1, 3, 3)), torch.from_numpy(b1), torch.from_numpy(gout)
        fns["torchvision-port"] = (lambda: ref_port.postprocessor_step(t[0], t[1], t[2], tw, tb, tg, True, 1.0),
                                   torch.get_num_threads())
    except ImportError:  # torchvision missing on this host: the C restatement alone is timed
        pass
    from oracle import c_oracle
    c_oracle.build()

    def c_step():
        c_oracle.forward(init, weight, offset, w9, b1, 1, 1.0)
        c_oracle.backward(gout, init, weight, offset, w9, 1, 1.0, need_grad_init=False)
    fns["c-oracle"] = (c_step, c_oracle.threads())
    return fns


def time_cpu(fn, warmup, steps):
    for _ in range(warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    return (time.perf_counter() - t0) / steps


def cpu_baseline(tiles, warmup=1, steps=2):
    detail = {}
    for name, (fn, cores) in cpu_step_fns(tiles).items():
        dt = time_cpu(fn, warmup, steps)
        detail[name] = {"value": tiles * TILE * TILE / dt / 1e9, "cores": cores}
    name = "torchvision-port" if "torchvision-port" in detail else "c-oracle"
    return {"value": detail[name]["value"], "unit": UNIT, "cores": detail[name]["cores"], "kind": "port",
            "sample": f"{tiles} tiles of {TILE}x{TILE} of the same workload (fwd+bwd, fp32) with {name} "
                      f"(the reference's call sites on torchvision's CPU operator); host has {os.cpu_count()} logical cores",
            "c_restatement": detail.get("c-oracle"), "all": detail}


def run_reference(args, rank, world):
    if rank != 0:
        return
    # bounded sample: size the per-step sample so a step takes about half a second on this host
    probe = cpu_step_fns(8)
    name = "torchvision-port" if "torchvision-port" in probe else "c-oracle"
    t8 = time_cpu(probe[name][0], 1, 1)
    tiles = int(min(max(4, round(8 * 0.5 / max(t8, 1e-6))), 4096, args.batch))
    fns = cpu_step_fns(tiles)
    fn, cores = fns[name]
    dt = time_cpu(fn, args.warmup, args.steps)
    v = tiles * TILE * TILE / dt / 1e9
    c_dt = time_cpu(fns["c-oracle"][0], 1, 2)
    sample = (f"each step = {tiles} tiles of {TILE}x{TILE} (bounded sample of the {args.batch}-tile workload), "
              f"fwd+bwd fp32 on host CPU with {name} ({cores} threads)")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "c_restatement": {"value": tiles * TILE * TILE / c_dt / 1e9,
                                               "cores": fns["c-oracle"][1]}},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def bind_to_gpu_cpus(torch, local_rank):
    """Pin this rank to the CPUs next to its GPU (NVML's ideal affinity), BEFORE any pinned host buffer exists, so that
    the e2e leg's host buffers are allocated on the GPU's own NUMA node and eight ranks do not pull their 7.8 GB per
    step across the socket interconnect.  Returns the number of CPUs bound to (0: left alone)."""
    if os.environ.get("JSPSR_BENCH_NO_AFFINITY", "0") == "1":
        return 0
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
            handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return len(os.sched_getaffinity(0))
    except Exception:
        return 0


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import jspsr_b200
    from jspsr_b200 import functional as F

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: jspsr_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    all_cpus = os.sched_getaffinity(0)
    cpus_bound = bind_to_gpu_cpus(torch, local_rank)
    if world > 1:
        # The gradient all-reduce is a 40-byte NCCL kernel enqueued behind compute kernels whose grids keep every SM full:
        # on a normal-priority stream it only gets a CTA slot when a compute grid drains, each rank at a different moment,
        # and the stream wait two steps later can stall.  High-priority NCCL streams are dispatched ahead of queued CTAs.
        os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
        dist.init_process_group("nccl", device_id=device)
    dtype = torch.float32 if args.dtype == "f32" else torch.bfloat16
    B = args.batch
    init, weight, offset, gout, w, b = make_inputs(torch, B, device, dtype, 1234 + rank)
    pp = jspsr_b200.PostProcessor(3, True, 1.0).to(device)
    with torch.no_grad():
        pp.w.copy_(w)
        pp.b.copy_(b)
    weight.requires_grad_(True)
    offset.requires_grad_(True)
    npix = B * TILE * TILE

    pending = []

    def step(ev=None):
        weight.grad = offset.grad = None
        pp.w.grad = pp.b.grad = None
        if ev:
            ev[0].record()
        out = pp(init, weight, offset)                 # 1 kernel
        if ev:
            ev[1].record()
        out.backward(gout)                             # 1 kernel
        if ev:
            ev[2].record()
        if world > 1:                                  # DDP's job for these two parameters (40 bytes): asynchronous on
            flat = torch.cat([pp.w.grad.reshape(-1), pp.b.grad.reshape(-1)])   # NCCL's stream, like a DDP bucket,
            pending.append((dist.all_reduce(flat, async_op=True), flat))       # so the next step's kernels are not held up
            while len(pending) > 4:
                pending.pop(0)[0].wait()
        return out

    def barrier():
        while pending:
            pending.pop(0)[0].wait()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = F.launch_count()
    sampler = ClockSampler(local_rank)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    t_start.record()
    for i in range(args.steps):
        step(evs[i])
    while pending:                       # the timed region ends when the last gradient all-reduce has landed
        pending.pop(0)[0].wait()
    t_end.record()
    barrier()
    clocks = sampler.stop()
    launches = F.launch_count() - launches0
    ms = t_start.elapsed_time(t_end) / args.steps
    fwd_ms = statistics.mean(e[0].elapsed_time(e[1]) for e in evs)
    bwd_ms = statistics.mean(e[1].elapsed_time(e[2]) for e in evs)
    if world > 1:
        t = torch.tensor([ms, fwd_ms, bwd_ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, fwd_ms, bwd_ms = t.tolist()
    value = world * npix / (ms * 1e-3) / 1e9

    peak, peak_src = peak_hbm()
    bwd_gbs = BWD_BYTES[args.dtype] * npix / (bwd_ms * 1e-3) / 1e9
    fwd_gbs = FWD_BYTES[args.dtype] * npix / (fwd_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "spn_backward_kernel", "achieved": bwd_gbs, "peak": peak, "unit": "GB/s",
                "frac": bwd_gbs / peak, "traffic": recorded_traffic("spn_backward_kernel"), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": BWD_BYTES[args.dtype] * npix, "avg_launch_ms": bwd_ms}
    roofline_fwd = {"bound": "hbm", "kernel": "spn_forward_kernel", "achieved": fwd_gbs, "peak": peak, "unit": "GB/s",
                    "frac": fwd_gbs / peak, "traffic": recorded_traffic("spn_forward_kernel"),
                    "algorithmic_bytes_per_launch": FWD_BYTES[args.dtype] * npix, "avg_launch_ms": fwd_ms,
                    "gpix_iter_per_s": npix / (fwd_ms * 1e-3) / 1e9}

    # ---- end to end: host (pinned) buffers -> module API -> host ----
    e2e = None
    if not args.no_e2e:
        host = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t.detach()) for t in (init, weight, offset, gout)]
        out_h = torch.empty(init.shape, dtype=dtype, pin_memory=True)
        gw_h = torch.empty(10, dtype=torch.float32, pin_memory=True)
        h2d = sum(t.numel() * t.element_size() for t in host)
        d2h = out_h.numel() * out_h.element_size() + gw_h.numel() * 4

        # The step is PCIe-bound (7.8 GB in): the batch goes through the module in chunks so that the H2D copy of chunk
        # i+1 (copy stream), fwd+bwd of chunk i (current stream) and the D2H of chunk i-1's output (third stream) overlap.
        n_chunks = 8 if B % 8 == 0 and B >= 64 else 1
        cb_ = B // n_chunks
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

        def e2e_step():
            cur = torch.cuda.current_stream()
            pp.w.grad = pp.b.grad = None
            s_in.wait_stream(cur)
            s_out.wait_stream(cur)

            def fetch(i):
                with torch.cuda.stream(s_in):
                    ts = [t[i * cb_:(i + 1) * cb_].to(device, non_blocking=True) for t in host]
                    ev = torch.cuda.Event()
                    ev.record(s_in)
                return ts, ev

            nxt = fetch(0)
            for i in range(n_chunks):
                (di, dw, do, dg), ev = nxt
                if i + 1 < n_chunks:
                    nxt = fetch(i + 1)
                cur.wait_event(ev)
                for t in (di, dw, do, dg):
                    t.record_stream(cur)
                dw.requires_grad_(True)
                do.requires_grad_(True)
                out = pp(di, dw, do)
                out.backward(dg)          # w.grad / b.grad accumulate over the chunks
                done = torch.cuda.Event()
                done.record(cur)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(done)
                    out.record_stream(s_out)
                    out_h[i * cb_:(i + 1) * cb_].copy_(out.detach(), non_blocking=True)
            flat = torch.cat([pp.w.grad.reshape(-1), pp.b.grad.reshape(-1)])
            if world > 1:
                dist.all_reduce(flat)
            gw_h.copy_(flat, non_blocking=True)
            cur.wait_stream(s_out)

        e2e_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.e2e_steps):
            e2e_step()
        e1.record()
        barrier()
        e_ms = e0.elapsed_time(e1) / args.e2e_steps
        if world > 1:
            t = torch.tensor([e_ms], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_ms = t.item()
        e2e = {"value": world * npix / (e_ms * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": e_ms,
               "host_cpus_bound": cpus_bound,
               "api": "jspsr_b200.PostProcessor.forward + backward on tensors copied from pinned host memory, "
                      f"{n_chunks} chunks (H2D / compute / D2H overlapped)"}
        del host, out_h

    extras = {}
    if not args.no_extras and rank == 0:
        extras = config_batch_latency(torch, jspsr_b200, F, device, dtype)
        if world == 1:
            del init, gout
            weight.grad = offset.grad = None
            torch.cuda.empty_cache()
            extras["variants"] = side_numbers(torch, F, device, min(B, 2048))

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic", "config": workload_config(args, world),
            "roofline": roofline, "roofline_fwd": roofline_fwd, "clocks": clocks, "e2e": e2e,
            "gpu_launches": launches, "tiles_per_s": value * 1e9 / (TILE * TILE), **extras}
    if rank == 0:
        if world == 1 and not args.no_cpu:
            os.sched_setaffinity(0, all_cpus)      # the CPU baseline uses every host core again
            torch.set_num_threads(len(all_cpus))
            line["cpu_baseline"] = cpu_baseline(args.cpu_tiles)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def side_numbers(torch, F, device, B):
    """Kernel-level numbers for the other variants of the path, same tile batch: bf16 I/O, the fixed-affinity
    T = 6 loop (NLSPN), and the backward that also returns grad_init.  Each with its algorithmic bytes."""
    peak, _ = peak_hbm()
    out = {}

    def timed(fn, n=5, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    npix = B * TILE * TILE
    init, weight, offset, gout, w, b = make_inputs(torch, B, device, torch.bfloat16, 4321)
    f = timed(lambda: F.spn_forward(init, weight, offset, w, b, 1, 1.0))
    g = timed(lambda: F.spn_backward(gout, init, weight, offset, w, 1, 1.0, need_grad_init=False))
    out["bf16_io"] = {"fwd_ms": f, "bwd_ms": g, "gpix_iter_per_s_fwd_bwd": npix / ((f + g) * 1e-3) / 1e9,
                      "fwd_frac_of_hbm_peak": FWD_BYTES["bf16"] * npix / (f * 1e-3) / 1e9 / peak,
                      "bwd_frac_of_hbm_peak": BWD_BYTES["bf16"] * npix / (g * 1e-3) / 1e9 / peak,
                      "note": "bf16 I/O, fp32 arithmetic; issue-bound (see profiles/r01_summary.md), not HBM-bound"}
    # torch.autocast(bfloat16) training: bf16 weight/offset from the Generator, fp32 DEM and grad_out (JSPSR_MIXED kernels)
    init32, gout32 = init.float(), gout.float()
    f = timed(lambda: F.spn_forward(init32, weight, offset, w, b, 1, 1.0))
    g = timed(lambda: F.spn_backward(gout32, init32, weight, offset, w, 1, 1.0, need_grad_init=False))
    out["autocast_bf16"] = {"fwd_ms": f, "bwd_ms": g, "gpix_iter_per_s_fwd_bwd": npix / ((f + g) * 1e-3) / 1e9,
                            "fwd_frac_of_hbm_peak": 62 * npix / (f * 1e-3) / 1e9 / peak,
                            "bwd_frac_of_hbm_peak": (8 + 54 + 54) * npix / (g * 1e-3) / 1e9 / peak,
                            "note": "bf16 weight/offset and their gradients, fp32 DEM / out / grad_out: 62 B/pixel forward, "
                                    "116 backward; issue-bound like the all-bf16 kernels"}
    del init, weight, offset, gout, init32, gout32
    init, weight, offset, gout, w, b = make_inputs(torch, B, device, torch.float32, 4322)
    g = timed(lambda: F.spn_backward(gout, init, weight, offset, w, 1, 1.0, need_grad_init=True))
    out["backward_with_grad_init"] = {"ms": g, "frac_of_hbm_peak": 228 * npix / (g * 1e-3) / 1e9 / peak,
                                      "note": "scatter into a block-floating-point tile of native integer shared-memory "
                                              "atomics (exact, order-independent inside a CTA); used by NLSPN only"}
    T = 6
    aff = weight * 0.1
    it = timed(lambda: F.spn_iterate(init, aff, offset, T), n=3)
    out["nlspn_loop_T6"] = {"ms": it, "gpix_iter_per_s": npix * T / (it * 1e-3) / 1e9,
                            "frac_of_hbm_peak_compulsory": (4 + 108 + 4 * T) * npix / (it * 1e-3) / 1e9 / peak,
                            "frac_of_hbm_peak_as_run": 116 * T * npix / (it * 1e-3) / 1e9 / peak,
                            "note": "T launches of the forward kernel, all T outputs kept"}
    del aff, gout, weight, offset
    torch.cuda.empty_cache()
    # the Generator's last two layers (1x1 convolutions C -> 9 / 16, sigmoid, zero centre
    # pair; spn.py:41-52,66-73) fused into the propagation forward; contraction on tcgen05 (3xTF32).  C = 128 is what
    # models/JSPSR.py builds at every YAML config (cat_only = True -> bc = num_feature = 32 -> bc * 4 channels).
    def gen_tail(C, Bg):
        npx = Bg * TILE * TILE
        g_ = torch.Generator(device=device).manual_seed(4323)
        ini = init[:Bg]
        feat = torch.randn(Bg, C, TILE, TILE, device=device, generator=g_)
        cw = 0.15 * torch.randn(25, C, device=device, generator=g_) * (64.0 / C) ** 0.5
        cw[9:] *= 1.3
        cb = 0.1 * torch.randn(25, device=device, generator=g_)
        fused = timed(lambda: F.gen_spn_forward(ini, feat, cw, cb, w, b, 1, 1.0, False))
        fused_wo = timed(lambda: F.gen_spn_forward(ini, feat, cw, cb, w, b, 1, 1.0, True))
        by = (C * 4 + 8) * npx
        res = {"C": C, "tiles": Bg, "ms": fused, "gpix_per_s": npx / (fused * 1e-3) / 1e9,
               "algorithmic_bytes_per_pixel": C * 4 + 8, "frac_of_hbm_peak": by / (fused * 1e-3) / 1e9 / peak,
               "ms_with_weight_offset_written": fused_wo,
               "frac_of_hbm_peak_with_weight_offset_written": (by + 108 * npx) / (fused_wo * 1e-3) / 1e9 / peak}
        if C != 128:
            return res
        cwt, cot = cw[:9].reshape(9, C, 1, 1).contiguous(), cw[9:].reshape(16, C, 1, 1).contiguous()

        def unfused():  # the reference's sequence (spn.py:66-73 + 99-118) with torch's convolutions and OUR propagation kernel
            weight = torch.sigmoid(torch.nn.functional.conv2d(feat, cwt, cb[:9]))
            o = torch.nn.functional.conv2d(feat, cot, cb[9:]).view(Bg, 8, 2, TILE, TILE)
            lo = list(torch.chunk(o, 8, dim=1))
            lo.insert(4, torch.zeros((Bg, 1, 2, TILE, TILE), device=device))
            return F.spn_forward(ini, weight, torch.cat(lo, dim=1).view(Bg, -1, TILE, TILE), w, b, 1, 1.0)

        un = timed(unfused, n=3)
        # training step through the fused tail: forward (weight/offset written) + spn_backward_kernel in GEN_PREACT mode
        # (writes the pre-activation gradients) + gen_grad_feature_kernel + the weight gradient as a library GEMM
        featg = feat.clone().requires_grad_()
        cwg, cbg = cw.clone().requires_grad_(), cb.clone().requires_grad_()
        wg, bg = w.clone().requires_grad_(), b.clone().requires_grad_()
        gout = torch.randn(Bg, 1, TILE, TILE, device=device, generator=g_)

        def train_fused():
            for t_ in (featg, cwg, cbg, wg, bg):
                t_.grad = None
            F.gen_propagate(ini, featg, cwg, cbg, wg, bg, 1, 1.0).backward(gout)

        cwt_g, cot_g = cwt.clone().requires_grad_(), cot.clone().requires_grad_()
        cbw_g, cbo_g = cb[:9].clone().requires_grad_(), cb[9:].clone().requires_grad_()

        def train_unfused():
            for t_ in (featg, cwt_g, cot_g, cbw_g, cbo_g, wg, bg):
                t_.grad = None
            weight = torch.sigmoid(torch.nn.functional.conv2d(featg, cwt_g, cbw_g))
            o = torch.nn.functional.conv2d(featg, cot_g, cbo_g).view(Bg, 8, 2, TILE, TILE)
            lo = list(torch.chunk(o, 8, dim=1))
            lo.insert(4, torch.zeros((Bg, 1, 2, TILE, TILE), device=device))
            F.propagate(ini, weight, torch.cat(lo, dim=1).view(Bg, -1, TILE, TILE), wg, bg, 1, 1.0).backward(gout)

        tr_f = timed(train_fused, n=3)
        tr_u = timed(train_unfused, n=2)
        gz = torch.randn(Bg, 25, TILE, TILE, device=device, generator=g_)
        gf = timed(lambda: F.gen_tail_grad_feature(gz, cw))
        del featg, gout, gz
        torch.cuda.empty_cache()
        feat16 = feat.bfloat16()
        fused16 = timed(lambda: F.gen_spn_forward(ini, feat16, cw, cb, w, b, 1, 1.0, False))
        fused16_wo = timed(lambda: F.gen_spn_forward(ini, feat16, cw, cb, w, b, 1, 1.0, True))
        res.update({
            "unfused_ms": un, "speedup_vs_unfused": un / fused,
            "training_step_ms": tr_f, "training_step_unfused_ms": tr_u, "training_speedup_vs_unfused": tr_u / tr_f,
            "grad_feature_kernel_ms": gf, "grad_feature_frac_of_hbm_peak": (100 + 4 * C) * npx / (gf * 1e-3) / 1e9 / peak,
            "autocast_bf16_features": {"ms": fused16, "frac_of_hbm_peak": (C * 2 + 8) * npx / (fused16 * 1e-3) / 1e9 / peak,
                                       "ms_with_weight_offset_written": fused16_wo,
                                       "frac_of_hbm_peak_with_weight_offset_written":
                                           (C * 2 + 8 + 54) * npx / (fused16_wo * 1e-3) / 1e9 / peak},
            "note": "gen_spn_forward_kernel: TMA ring -> tf32 hi/lo split into TMEM lanes -> tcgen05.mma (A from TMEM, "
                    "3-product split, fp32-level accuracy) -> per-pixel epilogue + 9-tap gather; unfused = torch 1x1 "
                    "convolutions (TF32 allowed, torch's default) + sigmoid + chunk/insert/cat + spn_forward_kernel"})
        return res

    out["generator_tail_fused"] = gen_tail(128, min(B, 1024))
    torch.cuda.empty_cache()
    out["generator_tail_fused"]["c64"] = gen_tail(64, min(B, 2048))
    torch.cuda.empty_cache()
    out["neighbours"] = neighbour_rows(torch, device, timed, peak, 4096)
    return out


def neighbour_rows(torch, device, timed, peak, B):
    """SURVEY.md section 8f ranks 3 and 4, the components either side of the propagation: the loss + gradient and the
    RMSE/MAE sums that read its output, the tile scheduler in front of it and the blended merge behind it.  Sizes far
    beyond L2; algorithmic bytes per pixel in the keys."""
    from jspsr_b200 import epilogue as EP, tiles as TL
    g = torch.Generator(device=device).manual_seed(99)
    gt = torch.rand(B, 1, TILE, TILE, device=device, generator=g)
    pred = gt + 0.05 * torch.randn(B, 1, TILE, TILE, device=device, generator=g)
    px = B * TILE * TILE
    res = {}
    t = timed(lambda: EP.loss_l1_l2_grad(pred, gt), n=20, warm=5)
    res["loss_l1_l2_sobel_with_gradient"] = {"ms": t, "bytes_per_pixel": 12, "frac_of_hbm_peak": 12 * px / (t * 1e-3) / 1e9 / peak,
                                            "tiles": B, "note": "issue-bound (two-level 3x3 stencil: Sobel, then its adjoint)"}
    t = timed(lambda: EP.loss_l1_l2_grad(pred, gt, want_grad=False), n=20, warm=5)
    res["loss_only"] = {"ms": t, "bytes_per_pixel": 8, "frac_of_hbm_peak": 8 * px / (t * 1e-3) / 1e9 / peak, "tiles": B}
    crop = int(TILE * 0.05)
    win = (TILE - 2 * crop) ** 2
    t = timed(lambda: EP.dem_metrics(pred, gt, 0.05, -80.0, 929.0, True), n=20, warm=5)
    res["rmse_mae_sums_log"] = {"ms": t, "bytes_per_window_pixel": 8, "frac_of_hbm_peak": 8 * B * win / (t * 1e-3) / 1e9 / peak,
                                "tiles": B, "note": "two expf per pixel: issue-bound"}
    del pred, gt
    k, stride, n = TILE, 103, 100
    side = stride * (n - 1) + k
    raster = torch.rand(1, side, side, device=device, generator=g)
    t = timed(lambda: TL.crop_tiles(raster, k, stride=stride, grid=(n, n)), n=20, warm=5)
    res["tile_crop"] = {"ms": t, "raster": [side, side], "tiles": n * n, "bytes_per_tile_pixel": 8,
                        "frac_of_hbm_peak": 8 * n * n * k * k / (t * 1e-3) / 1e9 / peak}
    tl = TL.crop_tiles(raster, k, stride=stride, grid=(n, n)).reshape(1, n * n, k, k)
    L = k - 2 * crop
    out_side = stride * (n - 1) + L
    for name, dt, ob in (("float64", torch.float64, 8), ("float32", torch.float32, 4)):
        t = timed(lambda: TL.merge_tiles(tl, 0.05, stride=stride, grid=(n, n), dtype=dt), n=20, warm=5)
        nbytes = 4 * n * n * L * L + ob * out_side * out_side
        res["blended_merge_" + name] = {"ms": t, "raster": [out_side, out_side], "tiles": n * n, "algorithmic_bytes": nbytes,
                                        "frac_of_hbm_peak": nbytes / (t * 1e-3) / 1e9 / peak}
    res["blended_merge_float64"]["note"] = "bit-identical to the reference's numpy float64 merge_dem (tests)"
    return res


def config_batch_latency(torch, jspsr_b200, F, device, dtype):
    """The YAML batch sizes (70 / 50 tiles; configs/*.yml:86) are launch-latency bound: report us per
    fwd+bwd with the two kernels captured in a CUDA graph, next to torchvision's CUDA operator when present."""
    res = {}
    for B in (70, 50, 2):
        init, weight, offset, gout, w, b = make_inputs(torch, B, device, dtype, 99)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                F.spn_forward(init, weight, offset, w, b, 1, 1.0)
                F.spn_backward(gout, init, weight, offset, w, 1, 1.0, need_grad_init=False)
        torch.cuda.current_stream().wait_stream(s)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            F.spn_forward(init, weight, offset, w, b, 1, 1.0)
            F.spn_backward(gout, init, weight, offset, w, 1, 1.0, need_grad_init=False)
        for _ in range(5):
            graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 200
        e0.record()
        for _ in range(n):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        res[f"B{B}"] = {"us_per_fwd_bwd": e0.elapsed_time(e1) / n * 1e3, "l2_resident": True}
        if dtype == torch.float32:
            # the training step as train/train_utils.py:205-214 chains it: propagation -> MultiLoss (L1 + L2 + 0.1 Grad,
            # losses and dTotal/dpred in one kernel) -> propagation backward, three launches in one graph
            from jspsr_b200 import epilogue as EP
            gt = (init + 0.05 * torch.randn_like(init)).clamp_(0, 1)
            with torch.cuda.stream(s):
                EP.loss_l1_l2_grad(init, gt)       # this stream's workspace is allocated outside the capture
            torch.cuda.current_stream().wait_stream(s)
            graph2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph2, stream=s):
                o = F.spn_forward(init, weight, offset, w, b, 1, 1.0)
                _, gl = EP.loss_l1_l2_grad(o, gt)
                F.spn_backward(gl, init, weight, offset, w, 1, 1.0, need_grad_init=False)
            for _ in range(5):
                graph2.replay()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(n):
                graph2.replay()
            e1.record()
            torch.cuda.synchronize()
            res[f"B{B}"]["us_per_fwd_loss_bwd"] = e0.elapsed_time(e1) / n * 1e3
            # the same step eager, through the drop-in modules and autograd (what an unmodified training loop pays):
            # wall clock per step, host-bound at these sizes (torch's own floor for a 3-node autograd step is ~135 us)
            post = jspsr_b200.PostProcessor(3, True, 1.0).to(device)
            crit = jspsr_b200.MultiLoss(L1=1.0, L2=1.0, Grad=0.1)
            wr, orq = weight.clone().requires_grad_(), offset.clone().requires_grad_()

            def eager_step():
                crit(post(init, wr, orq), gt)["Total"].backward()
                wr.grad = None
                orq.grad = None
                post.w.grad = None
                post.b.grad = None
            for _ in range(30):
                eager_step()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(n):
                eager_step()
            torch.cuda.synchronize()
            res[f"B{B}"]["eager_us_per_fwd_loss_bwd"] = (time.perf_counter() - t0) / n * 1e6
    out = {"config_batch": res}
    try:  # GPU incumbent: the unmodified call sequence on torchvision's CUDA kernels (a library), same shapes
        from oracle import ref_port
        B = 512
        init, weight, offset, gout, w, b = make_inputs(torch, B, device, torch.float32, 7)
        for _ in range(3):
            ref_port.postprocessor_step(init, weight, offset, w, b, gout, True, 1.0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ref_port.postprocessor_step(init, weight, offset, w, b, gout, True, 1.0)
        e1.record()
        torch.cuda.synchronize()
        out["gpu_incumbent_torchvision"] = {"value": B * TILE * TILE / (e0.elapsed_time(e1) / 5 * 1e-3) / 1e9,
                                            "unit": UNIT, "tiles": B, "note": "torchvision.ops.deform_conv2d CUDA "
                                            "+ 2 elementwise passes, fwd+bwd incl. grad_init (cannot be skipped there)"}
    except Exception as e:
        out["gpu_incumbent_torchvision"] = {"unavailable": str(e)[:200]}
    return out


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit(f"--gpus {args.gpus} needs torchrun (python -m torch.distributed.run --nproc-per-node {args.gpus} ...)")
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
