import os, sys
sys.path.insert(0, "/root/repo")
import torch
from jspsr_b200 import functional as F
def timed(fn, n=4, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for B in (2048, 70, 2):
    H = W = 128
    g = torch.Generator(device="cuda").manual_seed(3)
    feat = torch.rand(B, 1, H, W, device="cuda", generator=g)
    aff = 0.1 * torch.sigmoid(1.5 * torch.randn(B, 9, H, W, device="cuda", generator=g))
    off = (1.5 * torch.randn(B, 18, H, W, device="cuda", generator=g)).clamp_(-8, 8); off[:, 8:10] = 0
    for T in (1, 2, 3, 6, 8):
        gl = torch.randn(T, B, 1, H, W, device="cuda", generator=g)
        out = F.spn_iterate(feat, aff, off, T)
        def steps():
            carry = None; ga = go = None
            for t in range(T - 1, -1, -1):
                gg = gl[t] if carry is None else gl[t] + carry
                src = feat if t == 0 else out[t - 1]
                acc = None if ga is None else (ga, go)
                carry, ga, go, _, _ = F.spn_backward(gg, src, aff, off, None, 0, 0.0, need_grad_init=True, need_grad_w=False, accumulate_into=acc)
        a = timed(steps); b = timed(lambda: F.spn_iterate_backward(gl, feat, out, aff, off))
        print(f"B={B} T={T}: steps {a:8.3f} ms   split {b:8.3f} ms", flush=True)
