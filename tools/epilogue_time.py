"""Event-timed runs of the §8f rank 3/4 kernels at sizes far beyond L2 (B200): loss (+gradient), metrics,
tile crop, blended merge.  Prints ms, GB/s of algorithmic bytes and the fraction of MEASURED_PEAKS.json's HBM figure."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jspsr_b200  # noqa: E402
from jspsr_b200 import epilogue as EP, tiles as TL  # noqa: E402

PEAK = 6551.0
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timeit(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def report(name, ms, nbytes):
    gbs = nbytes / ms / 1e6
    print(f"{name:58s} {ms:8.3f} ms  {gbs:8.0f} GB/s  {gbs / PEAK:5.2f} of {PEAK:.0f}", flush=True)


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    for B in (4096, 70):
        gt = torch.rand(B, 1, 128, 128, device="cuda", generator=g)
        pred = gt + 0.05 * torch.randn(B, 1, 128, 128, device="cuda", generator=g)
        px = B * 128 * 128
        # the launches alone (outputs preallocated by the wrapper each call: allocation is a cache hit)
        report(f"loss L1+L2+Grad + gradient, {B} tiles (12 B/px)", timeit(lambda: EP.loss_l1_l2_grad(pred, gt)), 12 * px)
        report(f"loss only, {B} tiles (8 B/px)", timeit(lambda: EP.loss_l1_l2_grad(pred, gt, want_grad=False)), 8 * px)
        report(f"RMSE/MAE sums (log), {B} tiles (8 B/px on 116^2 of 128^2)",
               timeit(lambda: EP.dem_metrics(pred, gt, 0.05, -80.0, 929.0, True)), 8 * B * 116 * 116)
        if B == 70:
            # the unfused incumbent: torch ops of the reference's MultiLoss, forward + backward
            import torch.nn.functional as F
            kx = torch.tensor([[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]], device="cuda")
            k = (torch.stack([kx, kx.t()]) / 8.0)[:, None]

            def sg(x):
                return F.conv2d(F.pad(x, [1, 1, 1, 1], "replicate"), k)

            def unfused():
                p = pred.detach().requires_grad_()
                tot = F.l1_loss(p, gt) + F.mse_loss(p, gt) + 0.1 * F.l1_loss(sg(p), sg(gt))
                tot.backward()
                return p.grad
            report(f"  unfused torch MultiLoss fwd+bwd, {B} tiles", timeit(unfused), 12 * px)
    # tile scheduler / merge at raster scale: n x n tiles of 128, crop 6, stride 103
    for n in (100, 3):
        k, crop, stride = 128, 6, 103
        S = 1 if n > 3 else 1024
        side = stride * (n - 1) + k
        if n > 3:
            raster = torch.rand(1, side, side, device="cuda", generator=g)
            report(f"crop {n}x{n} tiles of 128 from {side}^2 (4+4 B per tile px)",
                   timeit(lambda: TL.crop_tiles(raster, k, stride=stride, grid=(n, n))), 8 * n * n * k * k)
        tiles = torch.rand(S, n * n, k, k, device="cuda", generator=g)
        L = k - 2 * crop
        out = stride * (n - 1) + L
        for dt, ob in ((torch.float64, 8), (torch.float32, 4)):
            report(f"merge {S} x {n}x{n} tiles -> {out}^2 {str(dt)[6:]} (4 B/used tile px + {ob} B/out px)",
                   timeit(lambda: TL.merge_tiles(tiles, 0.05, stride=stride, grid=(n, n), dtype=dt)),
                   S * (4 * n * n * L * L + ob * out * out))


if __name__ == "__main__":
    main()
