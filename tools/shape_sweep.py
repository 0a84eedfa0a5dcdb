"""Forward / backward fraction of the HBM peak across shapes (cliff detector): TMA and manual staging, aligned and
unaligned rows, tile batches and rasters, fp32 and autocast-mixed.  python tools/shape_sweep.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jspsr_b200 import functional as F
from tools.ab_hot import timeit

peak = 6551.4
g = torch.Generator(device="cuda").manual_seed(1)
w = torch.ones(1, 1, 3, 3, device="cuda"); b = torch.zeros(1, device="cuda")
shapes = [(2048, 128, 128), (600, 334, 334), (70, 128, 128), (64, 1000, 1000), (64, 1001, 1003), (4, 4096, 4096), (3, 5001, 5002),
          (1, 8192, 8192), (1, 2000, 30000), (20000, 32, 32), (512, 256, 64), (8, 3000, 200)]
for (B, H, W) in shapes:
    init = torch.rand(B, 1, H, W, device="cuda", generator=g)
    weight = torch.sigmoid(1.5 * torch.randn(B, 9, H, W, device="cuda", generator=g))
    offset = (1.5 * torch.randn(B, 18, H, W, device="cuda", generator=g)).clamp_(-8, 8)
    offset[:, 8:10] = 0
    gout = torch.randn(B, 1, H, W, device="cuda", generator=g)
    npx = B * H * W
    row = []
    for name, (w_, o_) in {"f32": (weight, offset), "mixed": (weight.bfloat16(), offset.bfloat16())}.items():
        f = timeit(lambda: F.spn_forward(init, w_, o_, w, b, 1, 1.0), n=7)
        bw = timeit(lambda: F.spn_backward(gout, init, w_, o_, w, 1, 1.0, need_grad_init=False), n=7)
        fb, bb = (116, 224) if name == "f32" else (62, 116)
        row.append(f"{name}: fwd {f:8.1f} us {fb * npx / f / 1e3 / peak:5.2f}  bwd {bw:8.1f} us {bb * npx / bw / 1e3 / peak:5.2f}")
    print(f"{B:6d} x {H:5d} x {W:5d} ({npx / 1e6:6.1f} Mpix)  " + " | ".join(row), flush=True)
    del init, weight, offset, gout
    torch.cuda.empty_cache()
