"""Tiny driver for ncu: python tools/prof_case.py <dtype f32|bf16|mixed> <B> <H> <W> [gi]   (mixed = autocast: bf16 weight / offset, fp32 DEM)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jspsr_b200 import functional as F
dt = torch.float32 if sys.argv[1] == "f32" else torch.bfloat16
dti = torch.float32 if sys.argv[1] in ("f32", "mixed") else torch.bfloat16
B, H, W = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
gi = len(sys.argv) > 5 and sys.argv[5] == "gi"
g = torch.Generator(device="cuda").manual_seed(1)
init = torch.rand(B, 1, H, W, device="cuda", generator=g).to(dti)
weight = torch.sigmoid(1.5 * torch.randn(B, 9, H, W, device="cuda", generator=g)).to(dt)
offset = (1.5 * torch.randn(B, 18, H, W, device="cuda", generator=g)).clamp_(-8, 8).to(dt)
gout = torch.randn(B, 1, H, W, device="cuda", generator=g).to(dti)
w = torch.ones(1, 1, 3, 3, device="cuda"); b = torch.zeros(1, device="cuda")
for _ in range(3):
    F.spn_forward(init, weight, offset, w, b, 1, 1.0)
    F.spn_backward(gout, init, weight, offset, w, 1, 1.0, need_grad_init=gi)
torch.cuda.synchronize()
print("done")
