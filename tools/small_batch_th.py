"""YAML batch sizes (70 / 50 / 2 tiles): forward / backward time against the rows-per-CTA choice.  python tools/small_batch_th.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jspsr_b200 import functional as F
from tools.ab_hot import timeit

g = torch.Generator(device="cuda").manual_seed(1)
w = torch.ones(1, 1, 3, 3, device="cuda"); b = torch.zeros(1, device="cuda")
for B in (70, 50, 16, 2):
    init = torch.rand(B, 1, 128, 128, device="cuda", generator=g)
    weight = torch.sigmoid(1.5 * torch.randn(B, 9, 128, 128, device="cuda", generator=g))
    offset = (1.5 * torch.randn(B, 18, 128, 128, device="cuda", generator=g)).clamp_(-8, 8)
    offset[:, 8:10] = 0
    gout = torch.randn(B, 1, 128, 128, device="cuda", generator=g)
    row = []
    for th in ("", "16", "8", "4", "2"):
        if th:
            os.environ["JSPSR_SPN_TILE_H"] = th
        else:
            os.environ.pop("JSPSR_SPN_TILE_H", None)
        f = timeit(lambda: F.spn_forward(init, weight, offset, w, b, 1, 1.0), n=31, warm=10)
        bw = timeit(lambda: F.spn_backward(gout, init, weight, offset, w, 1, 1.0, need_grad_init=False), n=31, warm=10)
        row.append(f"TH={th or 'auto':4s} fwd {f:6.1f} bwd {bw:6.1f}")
    os.environ.pop("JSPSR_SPN_TILE_H", None)
    print(f"B = {B:3d} (us)  " + " | ".join(row), flush=True)
