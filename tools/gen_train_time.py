"""Dev tool: time breakdown of a training step through the fused generator tail (fp32, 2048 tiles)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jspsr_b200 import functional as F
from tools.quick_bench import timeit
B, H, W, C = 2048, 128, 128, 64
init = torch.rand(B, 1, H, W, device="cuda"); feat = torch.randn(B, C, H, W, device="cuda", requires_grad=True)
cw = (torch.randn(25, C, device="cuda") * 0.15).requires_grad_(); cb = (torch.randn(25, device="cuda") * 0.1).requires_grad_()
w = torch.ones(1, 1, 3, 3, device="cuda", requires_grad=True); b = torch.zeros(1, device="cuda", requires_grad=True)
gout = torch.randn(B, 1, H, W, device="cuda")
def step():
    for t in (feat, cw, cb, w, b): t.grad = None
    out = F.gen_propagate(init, feat, cw, cb, w, b, 1, 1.0)
    out.backward(gout)
m, _ = timeit(step, n=5)
print(f"fused generator tail, training step (fwd + bwd): {m:.2f} ms")
out, weight, offset = F.gen_spn_forward(init, feat.detach(), cw.detach(), cb.detach(), w, b, 1, 1.0, True)
m, _ = timeit(lambda: F.gen_spn_forward(init, feat.detach(), cw.detach(), cb.detach(), w, b, 1, 1.0, True), n=5); print(f"  forward (weight/offset written) {m:.2f} ms")
m, _ = timeit(lambda: F.spn_backward(gout, init, weight, offset, w, 1, 1.0, need_grad_init=False), n=5); print(f"  spn_backward {m:.2f} ms")
gi, gwt, goff, gw, gb = F.spn_backward(gout, init, weight, offset, w, 1, 1.0, need_grad_init=False)
def make_gz():
    gz = torch.empty(B, 25, H * W, device="cuda"); gz4 = gz.view(B, 25, H, W)
    torch.mul(gwt, weight * (1.0 - weight), out=gz4[:, :9]); gz4[:, 9:17].copy_(goff[:, :8]); gz4[:, 17:].copy_(goff[:, 10:])
    return gz
m, _ = timeit(make_gz, n=5); print(f"  gz (sigmoid', offsets without the centre pair) {m:.2f} ms")
gz = make_gz()
m, _ = timeit(lambda: gz.sum(dim=2).sum(dim=0), n=5); print(f"  bias grad {m:.2f} ms")
fd = feat.detach().view(B, C, H * W)
m, _ = timeit(lambda: torch.bmm(gz, fd.transpose(1, 2)).sum(dim=0), n=5); print(f"  conv weight grad (bmm per sample) {m:.2f} ms")
cwd = cw.detach()
m, _ = timeit(lambda: F.gen_tail_grad_feature(gz.view(B, 25, H, W), cwd), n=5); print(f"  feature grad (tcgen05 kernel) {m:.2f} ms  ({B*H*W*356/m/1e6/6551.4:.3f} of HBM peak)")
m, _ = timeit(lambda: torch.matmul(cwd.t(), gz), n=5); print(f"  feature grad (matmul per sample) {m:.2f} ms")
# the unfused reference sequence, forward + backward, torch convs + our propagation
cwt = cw.detach()[:9].reshape(9, C, 1, 1).clone().requires_grad_(); cot = cw.detach()[9:].reshape(16, C, 1, 1).clone().requires_grad_()
cbw = cb.detach()[:9].clone().requires_grad_(); cbo = cb.detach()[9:].clone().requires_grad_()
def unfused():
    for t in (feat, cwt, cot, cbw, cbo, w, b): t.grad = None
    weight = torch.sigmoid(torch.nn.functional.conv2d(feat, cwt, cbw))
    o = torch.nn.functional.conv2d(feat, cot, cbo).view(B, 8, 2, H, W)
    lo = list(torch.chunk(o, 8, dim=1)); lo.insert(4, torch.zeros((B, 1, 2, H, W), device="cuda"))
    out = F.propagate(init, weight, torch.cat(lo, dim=1).view(B, -1, H, W), w, b, 1, 1.0)
    out.backward(gout)
m, _ = timeit(unfused, n=3); print(f"unfused (torch convs + our propagation), training step: {m:.2f} ms")
