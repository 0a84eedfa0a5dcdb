"""Dev tool: generator-tail kernel at C = 128 (EDSR / cat_only), fp32 and bf16 features."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jspsr_b200 import functional as F
from tools.quick_bench import timeit, PEAK
B, H, W, C = 1024, 128, 128, 128
init = torch.rand(B, 1, H, W, device="cuda"); feat = torch.randn(B, C, H, W, device="cuda")
cw = torch.randn(25, C, device="cuda") * 0.1; cb = torch.randn(25, device="cuda") * 0.1
w = torch.ones(1, 1, 3, 3, device="cuda"); b = torch.zeros(1, device="cuda")
npx = B * H * W
for f, es, tag in ((feat, 4, "fp32"), (feat.bfloat16(), 2, "bf16")):
    m, _ = timeit(lambda: F.gen_spn_forward(init, f, cw, cb, w, b, 1, 1.0, False))
    m2, _ = timeit(lambda: F.gen_spn_forward(init, f, cw, cb, w, b, 1, 1.0, True))
    print(f"C=128 {tag}: fused {m*1e3:.1f} us ({npx*(C*es+8)/m/1e6/PEAK:.3f})  with w/o written {m2*1e3:.1f} us ({npx*(C*es+8+27*es)/m2/1e6/PEAK:.3f})")
