"""Is the strip kernel slower in NORM_NONE mode than in NORM_RESIDUAL, or does the GPU slow down as the probe goes on?
Interleaves the two calls on the same 4096 x 32768 strip and reads SM clock / power between batches.
    python tools/strip_mode_probe.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pynvml
from jspsr_b200 import functional as F
import bench

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
dev = torch.device("cuda", 0)
rows, W = 4096, 32768
init, aff, off = bench.strip_rows(torch, dev, 0, rows, W, 11)
w = torch.full((1, 1, 3, 3), 1.05, device=dev); b = torch.full((1,), 0.1, device=dev)
status = torch.zeros(1, dtype=torch.int32, device=dev)
out = torch.empty_like(init)


def timed(fn, n=12):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    mhz = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
    watts = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, mhz, watts


def call(mode, ww, bb, a=aff):
    F.spn_forward_strip(init, a, off, ww, bb, mode, 1.0 if mode == 1 else 0.0, rows, 0, 0, status, out=out)

aff_small = aff * 0.1
for rep in range(4):
    for name, fn in (("residual w/b", lambda: call(1, w, b)), ("none no w/b", lambda: call(0, None, None)),
                     ("none w/b", lambda: call(0, w, b)), ("sum w/b", lambda: call(2, w, b)),
                     ("none, 0.1 x aff", lambda: call(0, None, None, aff_small))):
        ms, mhz, watts = timed(fn)
        print(f"rep {rep} {name:16s} {ms:.3f} ms   {mhz} MHz  {watts:.0f} W")
