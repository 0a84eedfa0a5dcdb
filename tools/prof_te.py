"""Tiny driver for ncu: the §8f rank 3/4 kernels at sizes far beyond L2.  python tools/prof_te.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jspsr_b200 import epilogue as EP, tiles as TL
g = torch.Generator(device="cuda").manual_seed(0)
B = 4096
gt = torch.rand(B, 1, 128, 128, device="cuda", generator=g)
pred = gt + 0.05 * torch.randn(B, 1, 128, 128, device="cuda", generator=g)
k, stride, n = 128, 103, 100
raster = torch.rand(1, stride * (n - 1) + k, stride * (n - 1) + k, device="cuda", generator=g)
for _ in range(3):
    EP.loss_l1_l2_grad(pred, gt)
    EP.dem_metrics(pred, gt, 0.05, -80.0, 929.0, True)
    t = TL.crop_tiles(raster, k, stride=stride, grid=(n, n))
    TL.merge_tiles(t.reshape(1, n * n, k, k), 0.05, stride=stride, grid=(n, n))
torch.cuda.synchronize()
print("done")
