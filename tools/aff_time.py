"""Dev tool: NLSPN affinity front-end timing (2048 tiles), every affinity flavour."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jspsr_b200 import functional as F
from tools.quick_bench import timeit, PEAK
B, H, W = 2048, 128, 128
g = torch.Generator(device="cuda").manual_seed(2)
conv_out = torch.randn(B, 24, H, W, device="cuda", generator=g)
conv_out[:, 16:] *= 60
conf = torch.rand(B, 1, H, W, device="cuda", generator=g)
gamma = torch.full((1,), 4.0, device="cuda")
go_ = torch.randn(B, 18, H, W, device="cuda", generator=g); ga_ = torch.randn(B, 9, H, W, device="cuda", generator=g)
npx = B * H * W
for aff in ("TGASS", "ASS", "AS", "TC"):
    fm, _ = timeit(lambda: F.nlspn_affinity_forward(conv_out, conf, gamma, aff))
    bm, _ = timeit(lambda: F.nlspn_affinity_backward(go_, ga_, conv_out, conf, gamma, aff))
    print(f"{aff:6s} fwd {fm*1e3:.1f} us ({npx*4*(24+1+27)/fm/1e6/PEAK:.3f}) | bwd {bm*1e3:.1f} us ({npx*4*(27+24+1+24+1)/bm/1e6/PEAK:.3f})", flush=True)
