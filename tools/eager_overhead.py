"""Host-side cost of one training step of the path at the YAML batch sizes, eager (no CUDA graph):
PostProcessor.forward -> MultiLoss -> backward through autograd.  Wall clock per step with the GPU kept busy
(no sync inside the loop) = max(host time, device time); also the device time alone from events."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jspsr_b200

for B in (70, 50, 2):
    g = torch.Generator(device="cuda").manual_seed(B)
    init = torch.rand(B, 1, 128, 128, device="cuda", generator=g)
    gt = (init + 0.05 * torch.randn(B, 1, 128, 128, device="cuda", generator=g)).clamp_(0, 1)
    weight = torch.sigmoid(1.5 * torch.randn(B, 9, 128, 128, device="cuda", generator=g)).requires_grad_()
    offset = (1.5 * torch.randn(B, 18, 128, 128, device="cuda", generator=g)).requires_grad_()
    post = jspsr_b200.PostProcessor(3, True, 1.0).cuda()
    crit = jspsr_b200.MultiLoss(L1=1.0, L2=1.0, Grad=0.1)

    def step():
        out = post(init, weight, offset)
        loss = crit(out, gt)["Total"]
        loss.backward()
        weight.grad = None; offset.grad = None; post.w.grad = None; post.b.grad = None

    for _ in range(20):
        step()
    torch.cuda.synchronize()
    n = 200
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    t_host = (time.perf_counter() - t0) / n * 1e6     # host enqueue time per step (GPU may lag behind)
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) / n * 1e6
    print(f"B={B:3d}: host enqueue {t_host:7.1f} us/step, wall incl. drain {t_all:7.1f} us/step", flush=True)
