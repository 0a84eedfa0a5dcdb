"""Host time per call (us) of the entry points at a tiny batch (B = 2: the kernels take a few us, so the loop is
host-bound and the numbers are host costs): raw forward through ctypes vs the C++ extension, the differentiable
module call, the loss, and torch's own floor for a three-node autograd step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jspsr_b200
from jspsr_b200 import _lib, functional as F, epilogue as EP


def per_call(fn, n=2000, warm=200):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


B = 2
g = torch.Generator(device="cuda").manual_seed(B)
init = torch.rand(B, 1, 128, 128, device="cuda", generator=g)
gt = init.clone()
weight = torch.sigmoid(torch.randn(B, 9, 128, 128, device="cuda", generator=g))
offset = torch.randn(B, 18, 128, 128, device="cuda", generator=g)
w = torch.ones(1, 1, 3, 3, device="cuda"); b = torch.zeros(1, device="cuda")
e = _lib.ext()
print("ctypes  spn_forward           ", round(per_call(lambda: F.spn_forward(init, weight, offset, w, b, 1, 1.0)), 1))
if e is not None:
    print("ext     spn_forward           ", round(per_call(lambda: e.spn_forward(init, weight, offset, w, b, 1, 1.0)), 1))
print("ctypes  loss (no grad)        ", round(per_call(lambda: EP.loss_l1_l2_grad(init, gt, want_grad=False)), 1))
post = jspsr_b200.PostProcessor(3, True, 1.0).cuda()
crit = jspsr_b200.MultiLoss(L1=1.0, L2=1.0, Grad=0.1)
with torch.no_grad():
    print("module  forward, no_grad      ", round(per_call(lambda: post(init, weight, offset)), 1))
    print("module  MultiLoss, no_grad    ", round(per_call(lambda: crit(init, gt)), 1))
wr, orq = weight.clone().requires_grad_(), offset.clone().requires_grad_()
print("module  forward, grad recorded", round(per_call(lambda: post(init, wr, orq)), 1))


def step():
    loss = crit(post(init, wr, orq), gt)["Total"]
    loss.backward()
    wr.grad = None; orq.grad = None; post.w.grad = None; post.b.grad = None


print("step    fwd + loss + backward ", round(per_call(step, n=500, warm=50), 1))


def torch_floor():   # three cheap differentiable torch ops + backward: what autograd itself costs
    loss = ((wr * 2.0).sum() + (orq * 2.0).sum()) * 0.5
    loss.backward()
    wr.grad = None; orq.grad = None


print("torch   3-op autograd floor   ", round(per_call(torch_floor, n=500, warm=50), 1))
