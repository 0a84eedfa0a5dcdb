"""Weight / bias gradient kernel of the Generator tail (gen_tail_wgrad.cu): accuracy against fp64 and time against the
library GEMM it replaces, for several accumulation-run lengths (JSPSR_GEN_WGRAD_RUN).
    python tools/gen_wgrad_probe.py [tiles] [C]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jspsr_b200 import functional as F

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
C = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = torch.device("cuda", 0)
torch.manual_seed(0)
gz = torch.randn(B, 25, 128, 128, device=dev)
feat = torch.randn(B, C, 128, 128, device=dev) + 0.5
peak = 6551.4


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


# fp64 reference on a subset of samples (exact enough: the full contraction in chunks)
ref_w = torch.zeros(25, C, dtype=torch.float64, device=dev)
ref_b = torch.zeros(25, dtype=torch.float64, device=dev)
for b0 in range(0, B, 64):
    g = gz[b0:b0 + 64].double().flatten(2)
    f = feat[b0:b0 + 64].double().flatten(2)
    ref_w += torch.bmm(g, f.transpose(1, 2)).sum(0)
    ref_b += g.sum((0, 2))
    del g, f
bytes_ = (25 + C) * 4 * B * 128 * 128
for run in (2, 4, 8, 16, 32, 64, 128, 512):
    os.environ["JSPSR_GEN_WGRAD_RUN"] = str(run)
    gw, gb = F.gen_tail_grad_params(gz, feat)
    err = float(((gw.double() - ref_w).abs().amax(1) / ref_w.abs().amax(1)).max())
    errb = float(((gb.double() - ref_b).abs() / ref_b.abs().max()).max())
    ms = timed(lambda: F.gen_tail_grad_params(gz, feat))
    print(f"run {run:4d}: {ms:.3f} ms = {bytes_ / ms / 1e6 / peak:.3f} of the HBM peak; max row-relative error w {err:.2e}  b {errb:.2e}")
os.environ.pop("JSPSR_GEN_WGRAD_RUN")
fv = feat.view(B, C, -1)
gv = gz.view(B, 25, -1)
lib = lambda: torch.bmm(gv, fv.transpose(1, 2)).sum(dim=0, dtype=torch.float32)
ms = timed(lib)
err = float(((lib().double() - ref_w).abs().amax(1) / ref_w.abs().amax(1)).max())
print(f"library bmm + sum (fp32 SIMT): {ms:.3f} ms = {bytes_ / ms / 1e6 / peak:.3f}; error {err:.2e}")
ms = timed(lambda: gv.sum(dim=2, dtype=torch.float32).sum(dim=0))
print(f"library bias reduction: {ms:.3f} ms")
