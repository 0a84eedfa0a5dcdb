"""Two-pixels-per-thread bf16 kernels (JSPSR_SPN_PAIR) against the one-pixel kernels: bit-identity, then same-box timing.
python tools/ab_pair.py [tiles]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jspsr_b200 import functional as F


def inputs(B, sigma, clip, seed=1):
    g = torch.Generator(device="cuda").manual_seed(seed)
    init = torch.rand(B, 1, 128, 128, device="cuda", generator=g)
    weight = torch.sigmoid(1.5 * torch.randn(B, 9, 128, 128, device="cuda", generator=g))
    offset = (sigma * torch.randn(B, 18, 128, 128, device="cuda", generator=g)).clamp_(-clip, clip)
    offset[:, 8:10] = 0
    gout = torch.randn(B, 1, 128, 128, device="cuda", generator=g)
    return init, weight, offset, gout


def variants(init, weight, offset, gout):
    return {"bf16": (init.bfloat16(), weight.bfloat16(), offset.bfloat16(), gout.bfloat16()),
            "mixed": (init, weight.bfloat16(), offset.bfloat16(), gout)}


def run(pair, fn):
    os.environ["JSPSR_SPN_PAIR"] = "1" if pair else "0"
    return fn()


def same(a, b):
    return bool(torch.equal(a.view(torch.int16 if a.dtype == torch.bfloat16 else torch.int32),
                            b.view(torch.int16 if b.dtype == torch.bfloat16 else torch.int32)))


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
    w = (torch.ones(1, 1, 3, 3, device="cuda") + 0.1 * torch.rand(1, 1, 3, 3, device="cuda")); b = torch.full((1,), 0.1, device="cuda")
    ok = True
    for B in (1, 3, 40, 300):
        for sigma, clip in ((1.5, 8.0), (16.0, 64.0)):
            for name, (i_, w_, o_, g_) in variants(*inputs(B, sigma, clip)).items():
                for mode in (0, 1, 2):
                    f0 = run(False, lambda: F.spn_forward(i_, w_, o_, w, b, mode, 0.7))
                    f1 = run(True, lambda: F.spn_forward(i_, w_, o_, w, b, mode, 0.7))
                    e = same(f0, f1)
                    b0 = run(False, lambda: F.spn_backward(g_, i_, w_, o_, w, mode, 0.7, need_grad_init=False))
                    b1 = run(True, lambda: F.spn_backward(g_, i_, w_, o_, w, mode, 0.7, need_grad_init=False))
                    # grad_weight / grad_offset: one owner per element (bit-identical); grad_w / grad_b: fp32 partial sums
                    # meet in fp64 atomics whose order is not fixed
                    eb = [same(x, y) if x.dim() == 4 and x.shape[1] > 1 else bool(torch.allclose(x, y, rtol=1e-5, atol=1e-6))
                          for x, y in zip(b0, b1) if x is not None]
                    if not (e and all(eb)):
                        ok = False
                        print("MISMATCH", B, sigma, name, mode, e, eb, flush=True)
    print("bit-identical forward / gradients:", ok, flush=True)
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    for name, (i_, w_, o_, g_) in variants(*inputs(B, 1.5, 8.0)).items():
        for rep in range(2):
            for pair in (False, True):
                tf = run(pair, lambda: timeit(lambda: F.spn_forward(i_, w_, o_, w, b, 1, 1.0)))
                tb = run(pair, lambda: timeit(lambda: F.spn_backward(g_, i_, w_, o_, w, 1, 1.0, need_grad_init=False)))
                print(f"{name} pair={int(pair)} fwd {tf:7.1f} us  bwd {tb:7.1f} us", flush=True)


if __name__ == "__main__":
    main()
