import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jspsr_b200 import functional as F
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from quick_bench import timeit, PEAK
g = torch.Generator(device="cuda").manual_seed(1)
for B, T in ((4096, 6), (1024, 6), (70, 6)):
    init = torch.rand(B, 1, 128, 128, device="cuda", generator=g)
    aff = torch.randn(B, 9, 128, 128, device="cuda", generator=g) * 0.1
    off = (1.5 * torch.randn(B, 18, 128, 128, device="cuda", generator=g)).clamp_(-8, 8)
    ref = None
    for mb in ("0", "16", "32", "48", "64", "96"):
        os.environ["JSPSR_SPN_ITER_CHUNK_MB"] = mb
        out = F.spn_iterate(init, aff, off, T)
        if ref is None: ref = out
        same = torch.equal(out, ref)
        m, best = timeit(lambda: F.spn_iterate(init, aff, off, T), n=5, warm=2)
        npx = B * 128 * 128
        comp = npx * (4 + 108 + 4 * T)
        print(f"B={B} T={T} chunk_MB={mb:>3s}: {m*1e3:9.1f} us  {npx*T/m/1e6:7.2f} Gpix.iter/s  compulsory-bytes rate {comp/m/1e6:6.0f} GB/s ({comp/m/1e6/PEAK:.3f})  bitwise==unblocked {same}", flush=True)
