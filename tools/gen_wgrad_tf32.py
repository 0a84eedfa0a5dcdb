"""Dev tool: the Generator tail's weight-gradient contraction ([B,25,HW] x [B,HW,C], K = pixels) - the one library GEMM
left in the fused training step - with torch's TF32 switch off (default) and on, C = 128, 1024 tiles; and the whole step.
The reference computes this gradient inside cuDNN's convolution backward, where TF32 is ON by default
(torch.backends.cudnn.allow_tf32); torch.bmm follows torch.backends.cuda.matmul.allow_tf32 (off by default)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jspsr_b200 import functional as F
from tools.quick_bench import timeit
B, H, W, C = 1024, 128, 128, 128
init = torch.rand(B, 1, H, W, device="cuda"); feat = torch.randn(B, C, H, W, device="cuda", requires_grad=True)
cw = (torch.randn(25, C, device="cuda") * 0.1).requires_grad_(); cb = (torch.randn(25, device="cuda") * 0.1).requires_grad_()
w = torch.ones(1, 1, 3, 3, device="cuda", requires_grad=True); b = torch.zeros(1, device="cuda", requires_grad=True)
gout = torch.randn(B, 1, H, W, device="cuda")
gz = torch.randn(B, 25, H * W, device="cuda") * 1e-3
fd = feat.detach().view(B, C, H * W)


def step():
    for t in (feat, cw, cb, w, b):
        t.grad = None
    F.gen_propagate(init, feat, cw, cb, w, b, 1, 1.0).backward(gout)


ref = torch.bmm(gz.double(), fd.double().transpose(1, 2)).sum(dim=0)
for tf32 in (False, True):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    m, _ = timeit(lambda: torch.bmm(gz, fd.transpose(1, 2)).sum(dim=0), n=5)
    got = torch.bmm(gz, fd.transpose(1, 2)).sum(dim=0)
    err = float((got.double() - ref).abs().max() / ref.abs().max())
    s, _ = timeit(step, n=5)
    print(f"allow_tf32={tf32}: weight-gradient bmm {m:.2f} ms ({B*H*W*(25+C)*4/m/1e6/6551.4:.2f} of HBM peak), "
          f"max error / max |value| {err:.1e}; fused training step {s:.2f} ms", flush=True)
