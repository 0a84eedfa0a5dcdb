"""Tiny driver for ncu: NLSPN affinity forward + backward, 2048 tiles."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jspsr_b200 import functional as F
B, H, W = 2048, 128, 128
g = torch.Generator(device="cuda").manual_seed(2)
conv_out = torch.randn(B, 24, H, W, device="cuda", generator=g)
conv_out[:, 16:] *= 60
conf = torch.rand(B, 1, H, W, device="cuda", generator=g)
gamma = torch.full((1,), 4.0, device="cuda")
go_ = torch.randn(B, 18, H, W, device="cuda", generator=g); ga_ = torch.randn(B, 9, H, W, device="cuda", generator=g)
for _ in range(3):
    F.nlspn_affinity_forward(conv_out, conf, gamma, "TGASS")
    F.nlspn_affinity_backward(go_, ga_, conv_out, conf, gamma, "TGASS")
torch.cuda.synchronize()
print("done")
