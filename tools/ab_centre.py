"""Same-box A/B of the centre-tap shortcut in the forward and backward kernels: the shortcut is taken when a warp's centre offsets are all
zero, so the same kernel is timed with offset[:, 8:10] = 0 (shortcut) and = 1e-20 (general path; identical positions except
on image row / column 0).  python tools/ab_centre.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jspsr_b200 import functional as F
from tools.ab_hot import timeit

g = torch.Generator(device="cuda").manual_seed(1)
B, H, W = 2048, 128, 128
init = torch.rand(B, 1, H, W, device="cuda", generator=g)
weight = torch.sigmoid(1.5 * torch.randn(B, 9, H, W, device="cuda", generator=g))
offset = (1.5 * torch.randn(B, 18, H, W, device="cuda", generator=g)).clamp_(-8, 8)
offset[:, 8:10] = 0
off_gen = offset.clone()
off_gen[:, 8:10] = 1e-20
w = torch.ones(1, 1, 3, 3, device="cuda") * 1.03; b = torch.full((1,), 0.1, device="cuda")
cases = {"f32": (init, weight, offset, off_gen), "mixed": (init, weight.bfloat16(), offset.bfloat16(), off_gen.bfloat16()),
         "bf16": (init.bfloat16(), weight.bfloat16(), offset.bfloat16(), off_gen.bfloat16())}
for rep in range(2):
    for name, (i_, w_, o0, o1) in cases.items():
        a = F.spn_forward(i_, w_, o0, w, b, 1, 0.9)
        c = F.spn_forward(i_, w_, o1, w, b, 1, 0.9)
        same = torch.equal(a[:, :, 1:, 1:], c[:, :, 1:, 1:])
        t0 = timeit(lambda: F.spn_forward(i_, w_, o0, w, b, 1, 1.0))
        t1 = timeit(lambda: F.spn_forward(i_, w_, o1, w, b, 1, 1.0))
        print(f"rep {rep} {name:6s} fwd shortcut {t0:7.1f} us   general {t1:7.1f} us   ratio {t1 / t0:.3f}   equal off row/col 0: {same}", flush=True)
        go = torch.randn(B, 1, H, W, device="cuda", generator=g).to(i_.dtype)
        ga = F.spn_backward(go, i_, w_, o0, w, 1, 0.9, need_grad_init=False)
        gc = F.spn_backward(go, i_, w_, o1, w, 1, 0.9, need_grad_init=False)
        same_b = all(torch.equal(a_[:, :, 1:, 1:], c_[:, :, 1:, 1:]) for a_, c_ in ((ga[1], gc[1]), (ga[2], gc[2])))
        t0 = timeit(lambda: F.spn_backward(go, i_, w_, o0, w, 1, 1.0, need_grad_init=False))
        t1 = timeit(lambda: F.spn_backward(go, i_, w_, o1, w, 1, 1.0, need_grad_init=False))
        print(f"rep {rep} {name:6s} bwd shortcut {t0:7.1f} us   general {t1:7.1f} us   ratio {t1 / t0:.3f}   equal off row/col 0: {same_b}", flush=True)
    aff = weight * 0.1
    t0 = timeit(lambda: F.spn_iterate(init, aff, offset, 6), n=5)
    t1 = timeit(lambda: F.spn_iterate(init, aff, off_gen, 6), n=5)
    print(f"rep {rep} loop T=6 shortcut {t0:7.1f} us   general {t1:7.1f} us   ratio {t1 / t0:.3f}", flush=True)
