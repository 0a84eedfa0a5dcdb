"""Why does one application of the fixed-affinity loop cost more than the T = 1 call on the same strip?  (bench.py `strips`,
N = 1: T6 = 1.09 x 6 x T1.)  Times the same strip kernel on a 4096 x 32768 strip while changing one thing at a time.
    python tools/strip_iter_probe.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jspsr_b200 import functional as F
import bench

dev = torch.device("cuda", 0)
rows, W = 4096, 32768
init, aff, off = bench.strip_rows(torch, dev, 0, rows, W, 11)
affn = aff / aff.sum(dim=1, keepdim=True)
w = torch.full((1, 1, 3, 3), 1.05, device=dev); b = torch.full((1,), 0.1, device=dev)
status = torch.zeros(1, dtype=torch.int32, device=dev)
A = init.clone(); Bb = torch.empty_like(init); pad = torch.empty(3 * 1024 * 1024 + 4096, device=dev)  # de-alias the next buffer
Cc = torch.empty_like(init)


def timed(fn, n=12):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def call(src, dst, a, mode, ww, bb):
    F.spn_forward_strip(src, a, off, ww, bb, mode, 1.0 if mode == 1 else 0.0, rows, 0, 0, status, out=dst)

state = {"i": 0}
def pingpong(a, x, y):
    s, d = (x, y) if state["i"] % 2 == 0 else (y, x)
    state["i"] += 1
    call(s, d, a, 0, None, None)

print("residual, w/b, A -> B            %.3f ms" % timed(lambda: call(A, Bb, aff, 1, w, b)))
print("none, no w/b, A -> B             %.3f ms" % timed(lambda: call(A, Bb, aff, 0, None, None)))
print("none, w/b, A -> B                %.3f ms" % timed(lambda: call(A, Bb, aff, 0, w, b)))
print("none, normalised aff, A -> B     %.3f ms" % timed(lambda: call(A, Bb, affn, 0, None, None)))
A.copy_(init)
print("none, normalised, ping-pong A<->B %.3f ms" % timed(lambda: pingpong(affn, A, Bb)))
A.copy_(init)
print("none, normalised, ping-pong A<->C (de-aliased) %.3f ms" % timed(lambda: pingpong(affn, A, Cc)))
A.copy_(init)
print("none, raw aff (values blow up), ping-pong A<->B %.3f ms" % timed(lambda: pingpong(aff, A, Bb), n=24))
print("finite after blow-up:", bool(torch.isfinite(A).all()), bool(torch.isfinite(Bb).all()))
