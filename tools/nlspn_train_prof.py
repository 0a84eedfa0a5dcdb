"""Dev tool: where the time goes in a training step through jspsr_b200.NLSPN (2048 tiles, T = 6, TGASS, conf_prop)."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import jspsr_b200
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
args = types.SimpleNamespace(prop_time=6, affinity="TGASS", affinity_gamma=0.5, conf_prop=True, preserve_input=False, legacy=False)
nl = jspsr_b200.NLSPN(args, 8, 1, 3, 3).cuda()
with torch.no_grad():
    nl.conv_offset_aff.weight.normal_(0, 0.3); nl.conv_offset_aff.bias.normal_(0, 0.5)
g = torch.randn(B, 8, 128, 128, device="cuda", requires_grad=True)
c = torch.rand(B, 1, 128, 128, device="cuda", requires_grad=True)
f0 = torch.rand(B, 1, 128, 128, device="cuda", requires_grad=True)
gout = torch.randn(B, 1, 128, 128, device="cuda")
def step():
    for t in (g, c, f0): t.grad = None
    nl.zero_grad()
    feat, lst, off, aff, gam = nl(f0, g, c)
    feat.backward(gout)
for _ in range(2): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
print(f"NLSPN training step, {B} tiles: {e0.elapsed_time(e1):.2f} ms")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=16, max_name_column_width=80))
