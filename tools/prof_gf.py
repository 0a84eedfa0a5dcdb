import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jspsr_b200 import functional as F
B, H, W, C = 2048, 128, 128, 64
gz = torch.randn(B, 25, H, W, device="cuda"); cw = torch.randn(25, C, device="cuda") * 0.15
for _ in range(3):
    F.gen_tail_grad_feature(gz, cw)
torch.cuda.synchronize(); print("done")
