"""Host vs device time of the fused Generator-tail training step (F.gen_propagate forward + backward) at the YAML
batch sizes, C = 128: wall clock per eager step (host-bound if it exceeds the event-timed device time)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jspsr_b200 import functional as F
C = 128
for B in (70, 50, 2):
    init = torch.rand(B, 1, 128, 128, device="cuda")
    feat = torch.randn(B, C, 128, 128, device="cuda").requires_grad_()
    cw = (torch.randn(25, C, device="cuda") * 0.1).requires_grad_(); cb = (torch.randn(25, device="cuda") * 0.1).requires_grad_()
    w = torch.ones(1, 1, 3, 3, device="cuda").requires_grad_(); b = torch.zeros(1, device="cuda").requires_grad_()
    gout = torch.randn(B, 1, 128, 128, device="cuda")

    def step():
        out = F.gen_propagate(init, feat, cw, cb, w, b, 1, 1.0)
        out.backward(gout)
        feat.grad = cw.grad = cb.grad = w.grad = b.grad = None

    for _ in range(20):
        step()
    torch.cuda.synchronize()
    n = 100
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    t_host = (time.perf_counter() - t0) / n * 1e6
    torch.cuda.synchronize()
    t_wall = (time.perf_counter() - t0) / n * 1e6
    print(f"B={B:3d}: host enqueue {t_host:7.1f} us/step, wall {t_wall:7.1f} us/step, device span {e0.elapsed_time(e1) / n * 1e3:7.1f} us/step", flush=True)
