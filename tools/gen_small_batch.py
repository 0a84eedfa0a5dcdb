"""Dev tool: generator-tail kernels at the YAML batch sizes (70 / 50 tiles per GPU, 1 in validation), C = 128."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jspsr_b200 import functional as F
from tools.quick_bench import timeit
C = 128
for B in (70, 50, 8, 1):
    init = torch.rand(B, 1, 128, 128, device="cuda"); feat = torch.randn(B, C, 128, 128, device="cuda")
    cw = torch.randn(25, C, device="cuda") * 0.1; cb = torch.randn(25, device="cuda") * 0.1
    w = torch.ones(1, 1, 3, 3, device="cuda"); b = torch.zeros(1, device="cuda")
    gz = torch.randn(B, 25, 128, 128, device="cuda")
    cwt, cot = cw[:9].reshape(9, C, 1, 1).contiguous(), cw[9:].reshape(16, C, 1, 1).contiguous()
    def unfused():
        weight = torch.sigmoid(torch.nn.functional.conv2d(feat, cwt, cb[:9]))
        o = torch.nn.functional.conv2d(feat, cot, cb[9:]).view(B, 8, 2, 128, 128)
        lo = list(torch.chunk(o, 8, dim=1)); lo.insert(4, torch.zeros((B, 1, 2, 128, 128), device="cuda"))
        return F.spn_forward(init, weight, torch.cat(lo, dim=1).view(B, -1, 128, 128), w, b, 1, 1.0)
    f, _ = timeit(lambda: F.gen_spn_forward(init, feat, cw, cb, w, b, 1, 1.0, False), n=30, warm=5)
    fw, _ = timeit(lambda: F.gen_spn_forward(init, feat, cw, cb, w, b, 1, 1.0, True), n=30, warm=5)
    gf, _ = timeit(lambda: F.gen_tail_grad_feature(gz, cw), n=30, warm=5)
    u, _ = timeit(unfused, n=10, warm=3)
    px = B * 16384
    print(f"B={B:3d}: fused fwd {f*1e3:7.1f} us ({px*520/f/1e6:6.0f} GB/s)  +w/o {fw*1e3:7.1f} us  grad_feature {gf*1e3:7.1f} us ({px*612/gf/1e6:6.0f} GB/s)  unfused fwd {u*1e3:7.1f} us", flush=True)
