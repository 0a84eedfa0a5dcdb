"""Tiny driver for ncu: python tools/prof_gen.py [B] [C]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jspsr_b200 import functional as F
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
H = W = 128; C = int(sys.argv[2]) if len(sys.argv) > 2 else 64
init = torch.rand(B, 1, H, W, device="cuda"); feat = torch.randn(B, C, H, W, device="cuda")
cw = torch.randn(25, C, device="cuda") * 0.15; cb = torch.randn(25, device="cuda") * 0.1
w = torch.ones(1, 1, 3, 3, device="cuda"); b = torch.zeros(1, device="cuda")
for _ in range(3):
    F.gen_spn_forward(init, feat, cw, cb, w, b, 1, 1.0, False)
torch.cuda.synchronize()
print("done")
