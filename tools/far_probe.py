"""Forward / backward time against the offset spread (how expensive is the out-of-tile global path?).  python tools/far_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jspsr_b200 import functional as F
from tools.ab_hot import timeit

g = torch.Generator(device="cuda").manual_seed(1)
B, H, W = 512, 128, 128
init = torch.rand(B, 1, H, W, device="cuda", generator=g)
weight = torch.sigmoid(1.5 * torch.randn(B, 9, H, W, device="cuda", generator=g))
base = torch.randn(B, 18, H, W, device="cuda", generator=g)
gout = torch.randn(B, 1, H, W, device="cuda", generator=g)
w = torch.ones(1, 1, 3, 3, device="cuda"); b = torch.zeros(1, device="cuda")
for sigma in (1.5, 4.0, 8.0, 16.0, 64.0):
    off = (sigma * base).clamp_(-4 * sigma, 4 * sigma)
    off[:, 8:10] = 0
    row = []
    for th in ("", "8", "4"):
        if th:
            os.environ["JSPSR_SPN_TILE_H"] = th
        else:
            os.environ.pop("JSPSR_SPN_TILE_H", None)
        f = timeit(lambda: F.spn_forward(init, weight, off, w, b, 1, 1.0), n=5)
        bw = timeit(lambda: F.spn_backward(gout, init, weight, off, w, 1, 1.0, need_grad_init=False), n=5)
        row.append(f"TH={th or 'auto'}: fwd {f:7.1f} bwd {bw:7.1f}")
    os.environ.pop("JSPSR_SPN_TILE_H", None)
    print(f"sigma {sigma:5.1f} us  " + " | ".join(row), flush=True)
