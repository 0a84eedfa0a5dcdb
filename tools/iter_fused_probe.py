"""Fixed-affinity loop (NLSPN): T launches vs the single-launch cluster kernel (JSPSR_SPN_ITER_FUSED=1).
    python tools/iter_fused_probe.py [tiles]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jspsr_b200 import functional as F
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
dev = torch.device("cuda", 0)
init, weight, offset, gout, w, b = bench.make_inputs(torch, B, dev, torch.float32, 4322)
aff = weight * 0.1
npix = B * 128 * 128
peak = 6551.4


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for T in (2, 6, 18):
    os.environ.pop("JSPSR_SPN_ITER_FUSED", None)
    ref = F.spn_iterate(init, aff, offset, T)
    t0 = timed(lambda: F.spn_iterate(init, aff, offset, T))
    os.environ["JSPSR_SPN_ITER_FUSED"] = "1"
    got = F.spn_iterate(init, aff, offset, T)
    t1 = timed(lambda: F.spn_iterate(init, aff, offset, T))
    comp = (4 + 108 + 4 * T) * npix
    print(f"T = {T:2d}: {T} launches {t0:.3f} ms ({comp / t0 / 1e6 / peak:.3f} of the HBM peak on compulsory bytes), "
          f"fused {t1:.3f} ms ({comp / t1 / 1e6 / peak:.3f}); per application {t0 / T:.3f} vs {t1 / T:.3f} ms; "
          f"bit-identical {bool(torch.equal(ref, got))}")
    del ref, got
