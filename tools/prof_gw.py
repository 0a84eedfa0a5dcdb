"""ncu target: the Generator tail's weight / bias gradient kernel (gen_tail_wgrad.cu) at C = 128, 1024 tiles."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jspsr_b200 import functional as F
B, H, W, C = 1024, 128, 128, 128
gz = torch.randn(B, 25, H, W, device="cuda"); feat = torch.randn(B, C, H, W, device="cuda")
for _ in range(3):
    F.gen_tail_grad_params(gz, feat)
torch.cuda.synchronize(); print("done")
