import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from jspsr_b200 import functional as F
B, H, W, C = 2048, 128, 128, 64
init = torch.rand(B, 1, H, W, device="cuda"); feat = torch.randn(B, C, H, W, device="cuda", requires_grad=True)
cw = (torch.randn(25, C, device="cuda") * 0.15).requires_grad_(); cb = (torch.randn(25, device="cuda") * 0.1).requires_grad_()
w = torch.ones(1, 1, 3, 3, device="cuda", requires_grad=True); b = torch.zeros(1, device="cuda", requires_grad=True)
gout = torch.randn(B, 1, H, W, device="cuda")
def step():
    for t in (feat, cw, cb, w, b): t.grad = None
    out = F.gen_propagate(init, feat, cw, cb, w, b, 1, 1.0)
    out.backward(gout)
for _ in range(2): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=90))
