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
cross-rank state of the path - grad_w[9] and grad_b[1] - is all-reduced INSIDE the backward kernel over peer memory
(NVLink stores + flags, include/jspsr_peer.h): no NCCL kernel, no extra launch and no stream wait in the step.
Every run also carries `ddp_parity` (N > 1: sharded gradients through that path against the single-process gradients
of the concatenated batch, checked before timing) and `strips` (BASELINE config 5: row-strip inference of a raster of
4096 x 32768 pixels per GPU, T = 1 and T = 6, halo exchange fused into the kernel, bit-compared with the unsharded
result first).

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
    ap.add_argument("--no-strips", action="store_true")
    ap.add_argument("--strip-rows", type=int, default=4096, help="rows of the 32768-wide raster per GPU (config 5)")
    return ap.parse_args()


def workload_config(args, n):
    return {
        "workload": f"configs/jspsr_r8_img.yml PostProcessor(3x3, residual) training step fwd+bwd, "
                    f"{args.batch} tiles of {TILE}x{TILE} per GPU, T=1, DEM detached",
        "tiles_per_gpu": args.batch, "tile": [TILE, TILE], "global_tiles": args.batch * n,
        "parallelism": f"dp{n} (batch-sharded tiles; grad_w/grad_b all-reduced inside the backward kernel over peer memory)",
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
            self._stop.wait(0.005)

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
    reducer = None
    ddp_parity = None
    if world > 1:
        # NCCL is plumbing only (barriers, gathering results, carrying the 64-byte IPC handles): the data path's one
        # exchange - 80 bytes of gradient sums per step - happens inside spn_backward_kernel over peer memory.
        dist.init_process_group("nccl", device_id=device)
        from jspsr_b200.peer import PeerGradReducer
        reducer = PeerGradReducer(average=True)
        ddp_parity = ddp_gradient_parity(torch, dist, jspsr_b200, reducer, device, rank, world)
    dtype = torch.float32 if args.dtype == "f32" else torch.bfloat16
    B = args.batch
    init, weight, offset, gout, w, b = make_inputs(torch, B, device, dtype, 1234 + rank)
    pp = jspsr_b200.PostProcessor(3, True, 1.0).to(device).set_grad_reducer(reducer)
    with torch.no_grad():
        pp.w.copy_(w)
        pp.b.copy_(b)
    weight.requires_grad_(True)
    offset.requires_grad_(True)
    npix = B * TILE * TILE

    def step(ev=None):
        weight.grad = offset.grad = None
        pp.w.grad = pp.b.grad = None
        if ev:
            ev[0].record()
        out = pp(init, weight, offset)                 # 1 kernel
        if ev:
            ev[1].record()
        out.backward(gout)                             # 1 kernel; N > 1: its last CTA all-reduces grad_w / grad_b with
        if ev:                                         # the other ranks' last CTAs (DDP's job for these two parameters),
            ev[2].record()                             # so pp.w.grad / pp.b.grad are the world's averages when it ends
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    out_keep = None
    for _ in range(max(args.warmup, 3)):
        out_keep = step()
    # everything slow on the host (NVML initialisation inside ClockSampler took 160 ms on one rank of four, event creation)
    # happens BEFORE the barrier: between the aligning step below and the timed loop there is only Python bookkeeping
    sampler = ClockSampler(local_rank)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if world > 1:
        # One more untimed step AFTER the host barrier: its backward exchanges gradients with every rank inside the
        # kernel, so the GPUs leave it within microseconds of each other with the timed steps already queued behind it.
        # Without it the ranks' host threads leave the barrier up to milliseconds apart, and the first timed step of the
        # early ranks measures that skew (it showed up as one 8.6 ms step among 3.9 ms ones at N = 8).
        out_keep = step()
    launches0 = F.launch_count()
    t_start.record()
    for i in range(args.steps):
        step(evs[i])
    t_end.record()
    # The clock / throttle sampler (NVML) starts once the K steps are ENQUEUED and samples while the GPU works through
    # them: an NVML query can hold the driver's per-device lock for milliseconds, and a kernel launch stuck behind it
    # showed up as one 11 ms step on one rank (N = 2, 10 steps: 4.46 instead of 3.71 ms per step).
    sampler.start()
    barrier()
    clocks = sampler.stop()
    launches = F.launch_count() - launches0
    ms = t_start.elapsed_time(t_end) / args.steps
    fwd_ms = statistics.mean(e[0].elapsed_time(e[1]) for e in evs)
    bwd_ms = statistics.mean(e[1].elapsed_time(e[2]) for e in evs)
    # per-step device times (start of step i -> start of step i + 1): a single hiccup and a steady gap look different
    marks = [e[0] for e in evs] + [t_end]
    per_step = [marks[i].elapsed_time(marks[i + 1]) for i in range(args.steps)]
    step_stats = {"p50_ms": statistics.median(per_step), "max_ms": max(per_step), "min_ms": min(per_step),
                  "slowest_step": per_step.index(max(per_step))}
    if world > 1:
        t = torch.tensor([ms, fwd_ms, bwd_ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, fwd_ms, bwd_ms = t.tolist()
        gathered = [None] * world
        dist.all_gather_object(gathered, step_stats)
        step_stats = {"per_rank": gathered, "p50_ms": max(g["p50_ms"] for g in gathered),
                      "max_ms": max(g["max_ms"] for g in gathered)}
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
            flat = torch.cat([pp.w.grad.reshape(-1), pp.b.grad.reshape(-1)])   # N > 1: already all-reduced by the kernels
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
        # what the host can deliver: a plain pinned cudaMemcpyAsync of 1 GiB per rank, all ranks at once (PCIe / NUMA
        # ceiling of this box at this N) - the e2e step is bound by it, not by the kernels
        probe_src = host[2].view(-1)[: (1 << 30) // host[2].element_size()]
        probe_dst = torch.empty_like(probe_src, device=device)
        probe_dst.copy_(probe_src, non_blocking=True)
        barrier()
        e0.record()
        for _ in range(3):
            probe_dst.copy_(probe_src, non_blocking=True)
        e1.record()
        barrier()
        probe_gbs = 3 * probe_src.numel() * probe_src.element_size() / (e0.elapsed_time(e1) * 1e-3) / 1e9
        h2d_gbs = h2d / (e_ms * 1e-3) / 1e9
        per_rank = [{"rank": rank, "h2d_probe_gbs": probe_gbs, "cpus_bound": cpus_bound}]
        if world > 1:
            gathered = [None] * world
            dist.all_gather_object(gathered, per_rank[0])
            per_rank = gathered
            probe_gbs = min(g["h2d_probe_gbs"] for g in gathered)
        del probe_dst
        e2e = {"value": world * npix / (e_ms * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": e_ms,
               "host_cpus_bound": cpus_bound,
               "h2d_gbs_per_rank": h2d_gbs, "h2d_probe_gbs_per_rank_min": probe_gbs,
               "frac_of_h2d_probe": h2d_gbs / probe_gbs, "per_rank": per_rank,
               "api": "jspsr_b200.PostProcessor.forward + backward on tensors copied from pinned host memory, "
                      f"{n_chunks} chunks (H2D / compute / D2H overlapped)"}
        del host, out_h

    extras = {}
    del init, gout, out_keep
    weight.grad = offset.grad = None
    del weight, offset
    torch.cuda.empty_cache()
    if not args.no_strips:
        # BASELINE config 5 (every rank takes part: the exchange is between neighbouring ranks' kernels)
        extras["strips"] = strip_inference(torch, dist, F, device, rank, world, args)
        torch.cuda.empty_cache()
    if ddp_parity is not None:
        extras["ddp_parity"] = ddp_parity
    extras["step_times"] = step_stats
    if not args.no_extras and rank == 0:
        extras.update(config_batch_latency(torch, jspsr_b200, F, device, dtype))
        if world == 1:
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
        reducer.close()
        dist.destroy_process_group()


def ddp_gradient_parity(torch, dist, jspsr_b200, reducer, device, rank, world):
    """DistributedDataParallel semantics on hardware, through the fused all-reduce: every rank runs PostProcessor
    forward + backward on its shard of a seeded batch (local mean loss, gradients of w / b averaged over the ranks inside
    the backward kernel) and compares with the single-process gradients of the concatenated batch (global mean loss),
    which every rank computes for itself.  Also checks that the reduced gradients are bit-identical on all ranks."""
    per, H, W = 6, 64, 128
    g = torch.Generator(device="cpu").manual_seed(20261018)
    n = per * world
    dem = torch.rand(n, 1, H, W, generator=g).to(device)
    weight = torch.sigmoid(1.5 * torch.randn(n, 9, H, W, generator=g)).to(device)
    offset = (1.5 * torch.randn(n, 18, H, W, generator=g)).to(device)
    gt = torch.rand(n, 1, H, W, generator=g).to(device)
    w0 = (1 + 0.2 * (torch.rand(1, 1, 3, 3, generator=g) - 0.5)).to(device)

    def run(sl, red):
        pp = jspsr_b200.PostProcessor(3, True, 1.0).to(device).set_grad_reducer(red)
        with torch.no_grad():
            pp.w.copy_(w0)
            pp.b.fill_(0.1)
        wt, of = weight[sl].clone().requires_grad_(), offset[sl].clone().requires_grad_()
        (pp(dem[sl], wt, of) - gt[sl]).square().mean().backward()
        return pp.w.grad.reshape(-1), pp.b.grad.reshape(-1), wt.grad, of.grad

    sl = slice(rank * per, (rank + 1) * per)
    gw, gb, gwt, gof = run(sl, reducer)                 # sharded, reduced inside the kernel
    rw, rb, rwt, rof = run(slice(0, n), None)           # single process, whole batch

    def rel(a, b):
        return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))

    flat = torch.cat([gw, gb])
    allg = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(allg, flat)
    res = torch.tensor([rel(gw, rw), rel(gb, rb), rel(gwt / world, rwt[sl]), rel(gof / world, rof[sl])],
                       device=device, dtype=torch.float64)
    dist.all_reduce(res, op=dist.ReduceOp.MAX)
    r = res.tolist()
    return {"max_rel": max(r), "grad_w_rel": r[0], "grad_b_rel": r[1], "grad_weight_rel": r[2], "grad_offset_rel": r[3],
            "bit_identical_across_ranks": all(torch.equal(a, flat) for a in allg), "tolerance": 2e-5,
            "ok": max(r) <= 2e-5, "tiles_per_rank": per, "tile": [H, W],
            "what": "PostProcessor fwd+bwd on rank shards (local mean loss), grad_w/grad_b averaged over ranks inside "
                    "spn_backward_kernel (peer memory), vs single-process gradients of the concatenated batch"}


def strip_rows(torch, device, r0, r1, W, seed, clip=6.0):
    """Rows [r0, r1) of a seeded synthetic raster (DEM, affinities, offsets): generated in 256-row blocks keyed by the
    global block index, so every sharding of the raster sees the same values."""
    blk = 256
    parts = ([], [], [])
    g = torch.Generator(device=device)
    for bi in range(r0 // blk, (r1 + blk - 1) // blk):
        g.manual_seed(seed * 100003 + bi)
        init = torch.rand(1, 1, blk, W, device=device, generator=g)
        aff = torch.sigmoid(1.5 * torch.randn(1, 9, blk, W, device=device, generator=g))
        off = (1.5 * torch.randn(1, 18, blk, W, device=device, generator=g)).clamp_(-clip, clip)
        off[:, 8:10] = 0
        lo, hi = max(r0, bi * blk) - bi * blk, min(r1, (bi + 1) * blk) - bi * blk
        for dst, t in zip(parts, (init, aff, off)):
            dst.append(t[:, :, lo:hi])
    return tuple(torch.cat(pr, dim=2).contiguous() for pr in parts)


def strip_inference(torch, dist, F, device, rank, world, args):
    """BASELINE config 5: a raster of (4096 * N) x 32768 pixels sharded into row strips, one per GPU (weak scaling: 4096 rows
    per rank).  T = 1 is JSPSR's single application, T = 6 the fixed-affinity loop; the halo exchange runs inside the
    propagation kernel (peer stores + flags, jspsr_b200/strips.py).  Before timing, a small raster is bit-compared with
    the unsharded single-GPU result on every rank."""
    from jspsr_b200.strips import StripPropagator, strip_bounds
    peak, _ = peak_hbm()
    w = torch.full((1, 1, 3, 3), 1.05, device=device)
    b = torch.full((1,), 0.1, device=device)
    halo = 8   # offsets are clipped to +-6 rows: ceil(6) + 2

    # ---- bit identity with the unsharded call (T = 1 and T = 3, all intermediates) ----
    Hs, Ws = 256 * world + 64, 512
    full = strip_rows(torch, device, 0, Hs, Ws, 7)
    ref1 = F.spn_forward(full[0], full[1], full[2], w, b, 1, 1.0)
    ref3 = F.spn_iterate(full[0], full[1] * 0.1, full[2], 3)
    r0, r1, _, _ = strip_bounds(Hs, world, rank, 0)
    band = [t[:, :, r0:r1].contiguous() for t in full]
    sp = StripPropagator(Hs, rank, world)
    ring = sp.peer_ring(r1 - r0, Ws, halo, n_buf=4)
    ring.load(band[0])
    out1, st = sp.forward_peer(ring, band[1], band[2], w, b, 1, 1.0)
    ok = torch.equal(out1, ref1[:, :, r0:r1])
    feats, st = sp.iterate_peer(ring, band[1] * 0.1, band[2], 3, keep_all=True)
    ok = ok and all(torch.equal(f, ref3[t][:, :, r0:r1]) for t, f in enumerate(feats))
    feats, st = sp.iterate_peer(ring, band[1] * 0.1, band[2], 2)       # continues from the ring's current band
    ref5 = F.spn_iterate(ref3[2], full[1] * 0.1, full[2], 2)
    ok = ok and torch.equal(feats[-1], ref5[1][:, :, r0:r1]) and int(st.item()) == 0
    flag = torch.tensor([1 if ok else 0], device=device)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    bitwise_ok = bool(flag.item())
    ring.close()
    del full, ref1, ref3, ref5, band, feats, out1, ring

    # ---- timing: 4096 x 32768 pixels per rank ----
    rows, W = args.strip_rows, 32768
    H_img = rows * world
    r0, r1 = rank * rows, (rank + 1) * rows
    init, aff, off = strip_rows(torch, device, r0, r1, W, 11)
    sp = StripPropagator(H_img, rank, world)
    ring = sp.peer_ring(rows, W, halo, n_buf=2)
    ring.load(init)
    del init
    out = torch.empty(1, 1, rows, W, device=device)
    res = {"raster": [H_img, W], "rows_per_gpu": rows, "halo_rows": halo, "bitwise_ok": bitwise_ok,
           "exchange": "none (one strip)" if world == 1 else
                       "fused into spn_forward_kernel: edge CTAs store their rows into the neighbours' next buffer over "
                       "NVLink and raise a flag; only the next application's edge CTAs wait for it",
           "halo_bytes_per_exchange_per_rank": 2 * halo * W * 4}

    def run(T, n):
        def one():
            if T == 1:
                sp.forward_peer(ring, aff, off, w, b, 1, 1.0, out=out)
            else:
                sp.iterate_peer(ring, aff, off, T)
        # Warm up for ~0.7 s of back-to-back launches first: under sustained load this GPU settles 10-15 % below its
        # 1965 MHz boost clock (power), and the strip kernel follows the SM clock (tools/strip_mode_probe.py: the same call
        # goes from 2.61 to 2.88 ms as the clock drops to 1700 MHz) - T = 1 and T = 6, and N = 1 and N = 8, are only
        # comparable when they are all measured in that state.
        t_warm = time.perf_counter()
        while True:
            for _ in range(8 if T == 1 else 2):
                one()
            torch.cuda.synchronize()
            done = torch.tensor([1.0 if time.perf_counter() - t_warm > 0.7 else 0.0], device=device)
            if world > 1:   # every rank must leave the loop after the same number of sequences (the protocol pairs them)
                dist.all_reduce(done, op=dist.ReduceOp.MIN)
            if done.item() > 0:
                break
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            one()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
        return {"ms": ms, "gpix": H_img * W * T / (ms * 1e-3) / 1e9,
                "frac_of_hbm_peak_per_gpu": 116 * rows * W * T / (ms * 1e-3) / 1e9 / peak, "bitwise_ok": bitwise_ok}

    res["T1"] = run(1, 24)
    res["T6"] = run(6, 4)
    res["measured"] = "after 0.7 s of back-to-back launches (sustained clocks), CUDA events, max over ranks"
    res["T6_over_6xT1"] = res["T6"]["ms"] / (6 * res["T1"]["ms"])
    res["status"] = int(ring.status.item())
    ring.close()
    return res


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
                                              "atomics (exact, order-independent inside a CTA); used when the propagated DEM carries a gradient (NLSPN "
                                              "loops the split backward does not take, callers that do not detach it)"}
    # SURVEY.md section 8d's "adversarial" set: offsets ~ N(0, 16^2), so most taps leave the staged tile and take the
    # bounds-checked global path (the price of unbounded offsets, not a configuration the reference produces)
    nf = min(B, 512)
    far = (16.0 * torch.randn((nf, 18, TILE, TILE), device=device,
                              generator=torch.Generator(device=device).manual_seed(77))).clamp_(-64, 64)
    far[:, 8:10] = 0
    f = timed(lambda: F.spn_forward(init[:nf], weight[:nf], far, w, b, 1, 1.0), n=3)
    g = timed(lambda: F.spn_backward(gout[:nf], init[:nf], weight[:nf], far, w, 1, 1.0, need_grad_init=False), n=3)
    pxf = nf * TILE * TILE
    out["far_offsets_sigma16"] = {"tiles": nf, "fwd_ms": f, "bwd_ms": g,
                                  "fwd_frac_of_hbm_peak": FWD_BYTES["f32"] * pxf / (f * 1e-3) / 1e9 / peak,
                                  "bwd_frac_of_hbm_peak": BWD_BYTES["f32"] * pxf / (g * 1e-3) / 1e9 / peak,
                                  "note": "offsets N(0, 16^2) clipped to +-64: taps outside the staged tile go through global "
                                          "loads with per-corner bounds checks (L2 gathers); results still exact"}
    del far
    # whole raster (runtime channel stride, narrow staged halo): one 8192 x 8192 image
    rs = 8192
    gr = torch.Generator(device=device).manual_seed(78)
    r_init = torch.rand(1, 1, rs, rs, device=device, generator=gr)
    r_w = torch.sigmoid(1.5 * torch.randn(1, 9, rs, rs, device=device, generator=gr))
    r_o = (1.5 * torch.randn(1, 18, rs, rs, device=device, generator=gr)).clamp_(-8, 8)
    r_g = torch.randn(1, 1, rs, rs, device=device, generator=gr)
    f = timed(lambda: F.spn_forward(r_init, r_w, r_o, w, b, 1, 1.0))
    g = timed(lambda: F.spn_backward(r_g, r_init, r_w, r_o, w, 1, 1.0, need_grad_init=False))
    out["raster_8192"] = {"fwd_ms": f, "bwd_ms": g, "fwd_frac_of_hbm_peak": FWD_BYTES["f32"] * rs * rs / (f * 1e-3) / 1e9 / peak,
                          "bwd_frac_of_hbm_peak": BWD_BYTES["f32"] * rs * rs / (g * 1e-3) / 1e9 / peak}
    del r_init, r_w, r_o, r_g
    torch.cuda.empty_cache()
    # SURVEY.md section 8d's largest single-image shape: 16384 x 16384 (31 GB of inputs, 29 GB of gradients), built
    # plane by plane so that the generator never holds a second copy
    if torch.cuda.mem_get_info(device)[0] > 90e9:
        try:
            rs = 16384
            r_init = torch.rand(1, 1, rs, rs, device=device, generator=gr)
            r_w = torch.empty(1, 9, rs, rs, device=device)
            r_o = torch.empty(1, 18, rs, rs, device=device)
            for k in range(9):
                r_w[0, k] = torch.sigmoid(1.5 * torch.randn(rs, rs, device=device, generator=gr))
            for k in range(18):
                r_o[0, k] = (1.5 * torch.randn(rs, rs, device=device, generator=gr)).clamp_(-8, 8)
            r_o[:, 8:10] = 0
            r_g = torch.randn(1, 1, rs, rs, device=device, generator=gr)
            f = timed(lambda: F.spn_forward(r_init, r_w, r_o, w, b, 1, 1.0), n=3, warm=1)
            g = timed(lambda: F.spn_backward(r_g, r_init, r_w, r_o, w, 1, 1.0, need_grad_init=False), n=3, warm=1)
            out["raster_16384"] = {"fwd_ms": f, "bwd_ms": g,
                                   "fwd_frac_of_hbm_peak": FWD_BYTES["f32"] * rs * rs / (f * 1e-3) / 1e9 / peak,
                                   "bwd_frac_of_hbm_peak": BWD_BYTES["f32"] * rs * rs / (g * 1e-3) / 1e9 / peak}
            del r_init, r_w, r_o, r_g
            torch.cuda.empty_cache()
        except RuntimeError as e:  # reported, never fatal for the bench line
            out["raster_16384"] = {"error": str(e)[:200]}
            torch.cuda.empty_cache()
    # SURVEY.md section 8d's T = 4: the LRRU-style cascade (LRRU.py:447-498) - four applications, each with fresh
    # weights / offsets, the previous output detached and blended with the valid input pixels by the model's own
    # torch code ((1 - mask) * out + mask * d_clear) in between.  Four quarter-batches stand in for the four stages.
    nq = B // 4
    if nq > 0:
        d_clear = init[:nq] * (torch.rand(nq, 1, TILE, TILE, device=device) > 0.9)

        def lrru(blend):
            o = init[:nq]
            for st in range(4):
                if blend == "torch":
                    mask = torch.sum(d_clear > 0.0, dim=1, keepdim=True)
                    mask = (mask > 0.0).type_as(d_clear)
                    o = (1.0 - mask) * o + mask * d_clear
                elif blend == "kernel":
                    o = F.preserve_blend(o, d_clear)
                o = F.spn_forward(o.detach(), weight[st * nq:(st + 1) * nq], offset[st * nq:(st + 1) * nq], w, b, 1, 1.0)
            return o
        t0 = timed(lambda: lrru(None), n=3)
        t1 = timed(lambda: lrru("torch"), n=3)
        t2 = timed(lambda: lrru("kernel"), n=3)
        out["lrru_cascade_T4"] = {"tiles": nq, "ms": t0, "gpix_iter_per_s": 4 * nq * TILE * TILE / (t0 * 1e-3) / 1e9,
                                  "frac_of_hbm_peak": 4 * FWD_BYTES["f32"] * nq * TILE * TILE / (t0 * 1e-3) / 1e9 / peak,
                                  "ms_with_the_models_torch_blend": t1, "ms_with_preserve_blend_kernel": t2,
                                  "note": "4 forward launches with fresh weights (Post_process_deconv x 4); the blend between "
                                          "stages is the reference model's own elementwise torch code (six launches), outside "
                                          "the module; jspsr_preserve_blend does the same arithmetic in one pass"}
        del d_clear
    T = 6
    aff = weight * 0.1
    it = timed(lambda: F.spn_iterate(init, aff, offset, T), n=3)
    out["nlspn_loop_T6"] = {"ms": it, "gpix_iter_per_s": npix * T / (it * 1e-3) / 1e9,
                            "frac_of_hbm_peak_compulsory": (4 + 108 + 4 * T) * npix / (it * 1e-3) / 1e9 / peak,
                            "frac_of_hbm_peak_as_run": 116 * T * npix / (it * 1e-3) / 1e9 / peak,
                            "note": "T launches of the forward kernel, all T outputs kept"}
    # the single-launch form (one 16-CTA cluster per sample, tap state in registers, feature in distributed shared memory)
    os.environ["JSPSR_SPN_ITER_FUSED"] = "1"
    try:
        fused = F.spn_iterate(init, aff, offset, T)
        itf = timed(lambda: F.spn_iterate(init, aff, offset, T), n=3)
    finally:
        os.environ.pop("JSPSR_SPN_ITER_FUSED", None)
    same = bool(torch.equal(fused, F.spn_iterate(init, aff, offset, T)))
    del fused
    out["nlspn_loop_T6"]["fused_single_launch"] = {
        "ms": itf, "bit_identical_to_T_launches": same,
        "frac_of_hbm_peak_compulsory": (4 + 108 + 4 * T) * npix / (itf * 1e-3) / 1e9 / peak,
        "note": "JSPSR_SPN_ITER_FUSED=1 (spn_iterate_fused.cu); the default is whichever of the two is faster here"}
    # the loop as a training step (CompletionFormer-style: loss on the last step's output), forward + backward through
    # autograd: T applications of the full backward with accumulation ("steps") against T light carry launches + one
    # gradient kernel that sums over t in registers ("split", the default; spn_iterate_backward.cu)
    fa, aa, oa = init.clone().requires_grad_(True), aff.clone().requires_grad_(True), offset.clone().requires_grad_(True)

    def train_step():
        F.iterate(fa, aa, oa, T)[-1].backward(gout)
        fa.grad = aa.grad = oa.grad = None
    step_ms = {}
    for mode in ("steps", "split"):
        os.environ["JSPSR_ITER_BWD"] = mode
        try:
            step_ms[mode] = timed(train_step, n=3)
        finally:
            os.environ.pop("JSPSR_ITER_BWD", None)
    out["nlspn_loop_T6"]["train_step"] = {
        "fwd_bwd_ms": step_ms["split"], "fwd_bwd_ms_T_backward_applications": step_ms["steps"],
        "bwd_ms": step_ms["split"] - it, "bwd_ms_T_backward_applications": step_ms["steps"] - it,
        "note": "backward of the loop: aff / offset are the same at every step, so the step-to-step gradient runs as T "
                "scatter-only launches and the 27 gradients are summed over t in registers against the T staged features "
                "(jspsr_spn_iterate_backward); JSPSR_ITER_BWD=steps keeps T full backward applications, whose "
                "read-modify-write of 27 channels per step is HBM-bound"}
    del fa, aa, oa
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
        # (writes the pre-activation gradients) + gen_grad_feature_kernel + gen_grad_weight_kernel (no library GEMM left)
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
        gp = timed(lambda: F.gen_tail_grad_params(gz, feat))          # weight + bias gradients, one pass (tcgen05)
        F._GEN_WGRAD_LIBRARY = True                                   # the cuBLAS bmm + reduction the kernel replaced
        tr_lib = timed(train_fused, n=3)
        F._GEN_WGRAD_LIBRARY = False
        del featg, gout, gz
        torch.cuda.empty_cache()
        feat16 = feat.bfloat16()
        fused16 = timed(lambda: F.gen_spn_forward(ini, feat16, cw, cb, w, b, 1, 1.0, False))
        fused16_wo = timed(lambda: F.gen_spn_forward(ini, feat16, cw, cb, w, b, 1, 1.0, True))
        gz16 = torch.randn(Bg, 25, TILE, TILE, device=device, generator=g_).bfloat16()
        gp16 = timed(lambda: F.gen_tail_grad_params(gz16, feat16))
        gf16 = timed(lambda: F.gen_tail_grad_feature(gz16, cw))
        del gz16
        res.update({
            "unfused_ms": un, "speedup_vs_unfused": un / fused,
            "training_step_ms": tr_f, "training_step_unfused_ms": tr_u, "training_speedup_vs_unfused": tr_u / tr_f,
            "grad_feature_kernel_ms": gf, "grad_feature_frac_of_hbm_peak": (100 + 4 * C) * npx / (gf * 1e-3) / 1e9 / peak,
            "grad_params_kernel_ms": gp, "grad_params_frac_of_hbm_peak": (100 + 4 * C) * npx / (gp * 1e-3) / 1e9 / peak,
            "training_step_ms_with_library_weight_gradient": tr_lib,
            "autocast_bf16_features": {"ms": fused16, "frac_of_hbm_peak": (C * 2 + 8) * npx / (fused16 * 1e-3) / 1e9 / peak,
                                       "grad_params_kernel_ms": gp16,
                                       "grad_params_frac_of_hbm_peak": (50 + 2 * C) * npx / (gp16 * 1e-3) / 1e9 / peak,
                                       "grad_feature_kernel_ms": gf16,
                                       "grad_feature_frac_of_hbm_peak": (50 + 2 * C) * npx / (gf16 * 1e-3) / 1e9 / peak,
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
                                "tiles": B, "note": "float4 kernel (91 us alone, 5.3 TB/s of DRAM reads: whole 32-byte sectors of the window rows) + memset + two small torch launches for mean / sqrt"}
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
        if dtype == torch.float32 and B == 70:
            # BASELINE config 2 at its own batch: the tensors torch.autocast(bfloat16) hands to the layer (bf16 weight /
            # offset from the Generator's convolutions, fp32 DEM), same two launches in a graph
            wb, ob = weight.detach().bfloat16(), offset.detach().bfloat16()
            with torch.cuda.stream(s):
                for _ in range(3):
                    F.spn_forward(init, wb, ob, w, b, 1, 1.0)
                    F.spn_backward(gout, init, wb, ob, w, 1, 1.0, need_grad_init=False)
            torch.cuda.current_stream().wait_stream(s)
            graph_a = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph_a, stream=s):
                F.spn_forward(init, wb, ob, w, b, 1, 1.0)
                F.spn_backward(gout, init, wb, ob, w, 1, 1.0, need_grad_init=False)
            for _ in range(5):
                graph_a.replay()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(n):
                graph_a.replay()
            e1.record()
            torch.cuda.synchronize()
            res[f"B{B}"]["autocast_bf16_us_per_fwd_bwd"] = e0.elapsed_time(e1) / n * 1e3
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
