"""Same-box A/B of the hot kernels: python tools/ab_hot.py <tag>.  Prints median-of-15 event times (us)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jspsr_b200 import functional as F


def timeit(fn, n=15, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    return t[len(t) // 2] * 1e3


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else ""
    g = torch.Generator(device="cuda").manual_seed(1)
    B, H, W = 2048, 128, 128
    init = torch.rand(B, 1, H, W, device="cuda", generator=g)
    weight = torch.sigmoid(1.5 * torch.randn(B, 9, H, W, device="cuda", generator=g))
    offset = (1.5 * torch.randn(B, 18, H, W, device="cuda", generator=g)).clamp_(-8, 8)
    offset[:, 8:10] = 0
    gout = torch.randn(B, 1, H, W, device="cuda", generator=g)
    w = torch.ones(1, 1, 3, 3, device="cuda"); b = torch.zeros(1, device="cuda")
    res = {}
    for name, (i_, w_, o_, g_) in {"f32": (init, weight, offset, gout),
                                   "bf16": (init.bfloat16(), weight.bfloat16(), offset.bfloat16(), gout.bfloat16()),
                                   "mixed": (init, weight.bfloat16(), offset.bfloat16(), gout)}.items():
        res[name + " fwd"] = timeit(lambda: F.spn_forward(i_, w_, o_, w, b, 1, 1.0))
        res[name + " bwd"] = timeit(lambda: F.spn_backward(g_, i_, w_, o_, w, 1, 1.0, need_grad_init=False))
    res["f32 bwd+gi"] = timeit(lambda: F.spn_backward(gout, init, weight, offset, w, 1, 1.0, need_grad_init=True))
    conv_out = torch.randn(B, 24, H, W, device="cuda", generator=g)
    conv_out[:, 16:] *= 60
    conf = torch.rand(B, 1, H, W, device="cuda", generator=g)
    gamma = torch.full((1,), 4.0, device="cuda")
    go_ = torch.randn(B, 18, H, W, device="cuda", generator=g); ga_ = torch.randn(B, 9, H, W, device="cuda", generator=g)
    res["aff fwd"] = timeit(lambda: F.nlspn_affinity_forward(conv_out, conf, gamma, "TGASS"))
    res["aff bwd"] = timeit(lambda: F.nlspn_affinity_backward(go_, ga_, conv_out, conf, gamma, "TGASS"))
    print(tag, " | ".join(f"{k} {v:7.1f}" for k, v in res.items()), flush=True)


if __name__ == "__main__":
    main()
