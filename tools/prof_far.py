"""ncu target: forward / backward with offsets N(0, 16^2) (most taps leave the staged tile), 512 tiles."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jspsr_b200 import functional as F
g = torch.Generator(device="cuda").manual_seed(1)
B, H, W = 512, 128, 128
init = torch.rand(B, 1, H, W, device="cuda", generator=g)
weight = torch.sigmoid(1.5 * torch.randn(B, 9, H, W, device="cuda", generator=g))
off = (16.0 * torch.randn(B, 18, H, W, device="cuda", generator=g)).clamp_(-64, 64)
off[:, 8:10] = 0
gout = torch.randn(B, 1, H, W, device="cuda", generator=g)
w = torch.ones(1, 1, 3, 3, device="cuda"); b = torch.zeros(1, device="cuda")
for _ in range(3):
    F.spn_forward(init, weight, off, w, b, 1, 1.0)
    F.spn_backward(gout, init, weight, off, w, 1, 1.0, need_grad_init=False)
torch.cuda.synchronize(); print("done")
