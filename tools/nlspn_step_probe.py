"""NLSPN fixed-affinity loop, training step (T = 6): forward, backward, and the backward's pieces (dev tool).
python tools/nlspn_step_probe.py [tiles]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jspsr_b200 import functional as F


def timed(fn, n=5, warm=2):
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


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    T, H, W = 6, 128, 128
    g = torch.Generator(device="cuda").manual_seed(3)
    feat = torch.rand(B, 1, H, W, device="cuda", generator=g)
    aff = 0.1 * torch.sigmoid(1.5 * torch.randn(B, 9, H, W, device="cuda", generator=g))
    off = (1.5 * torch.randn(B, 18, H, W, device="cuda", generator=g)).clamp_(-8, 8)
    off[:, 8:10] = 0
    gout = torch.randn(B, 1, H, W, device="cuda", generator=g)
    npx = B * H * W
    print(f"{B} tiles, T = {T}")
    t = timed(lambda: F.spn_iterate(feat, aff, off, T))
    print(f"forward loop           {t:7.3f} ms  ({116 * T * npx / t / 1e6:6.0f} GB/s as run)")
    t = timed(lambda: F.spn_backward(gout, feat, aff, off, None, 0, 0.0, need_grad_init=True, need_grad_w=False))
    print(f"backward step, write   {t:7.3f} ms  ({228 * npx / t / 1e6:6.0f} GB/s at 228 B/pixel)")
    _, ga, go, _, _ = F.spn_backward(gout, feat, aff, off, None, 0, 0.0, need_grad_init=True, need_grad_w=False)
    t = timed(lambda: F.spn_backward(gout, feat, aff, off, None, 0, 0.0, need_grad_init=True, need_grad_w=False,
                                     accumulate_into=(ga, go)))
    print(f"backward step, ACC     {t:7.3f} ms  ({228 * npx / t / 1e6:6.0f} GB/s at 228 B/pixel; the REDs read and write their 108 B)")
    fa = feat.clone().requires_grad_(True)
    aa = aff.clone().requires_grad_(True)
    oa = off.clone().requires_grad_(True)

    def step():
        out = F.iterate(fa, aa, oa, T)
        out[-1].backward(gout)
        fa.grad = aa.grad = oa.grad = None
    for mode in ("steps", "split"):
        os.environ["JSPSR_ITER_BWD"] = mode
        t = timed(step, n=3)
        print(f"autograd fwd + bwd     {t:7.3f} ms   (JSPSR_ITER_BWD={mode})")
    gl = torch.zeros(T, B, 1, H, W, device="cuda")
    gl[-1] = gout
    out = F.spn_iterate(feat, aff, off, T)
    t = timed(lambda: F.spn_iterate_backward(gl, feat, out, aff, off))
    print(f"split backward alone   {t:7.3f} ms")


if __name__ == "__main__":
    main()
