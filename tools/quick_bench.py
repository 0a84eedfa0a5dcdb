"""Kernel-level timing sweep (dev tool): fwd / bwd at several shapes and dtypes, CUDA events, GB/s vs measured peak."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jspsr_b200 import functional as F

PEAK = 6551.4
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    return t[len(t) // 2], t[0]


def run(B, H, W, dtype, sigma=1.5, mode=1, gi=False, tag=""):
    g = torch.Generator(device="cuda").manual_seed(1)
    init = torch.rand(B, 1, H, W, device="cuda", generator=g).to(dtype)
    weight = torch.sigmoid(1.5 * torch.randn(B, 9, H, W, device="cuda", generator=g)).to(dtype)
    offset = (sigma * torch.randn(B, 18, H, W, device="cuda", generator=g)).clamp_(-8 * max(1, sigma / 1.5), 8 * max(1, sigma / 1.5))
    offset[:, 8:10] = 0
    offset = offset.to(dtype)
    gout = torch.randn(B, 1, H, W, device="cuda", generator=g).to(dtype)
    w = torch.ones(1, 1, 3, 3, device="cuda"); b = torch.zeros(1, device="cuda")
    es = 4 if dtype == torch.float32 else 2
    npx = B * H * W
    fm, fb = timeit(lambda: F.spn_forward(init, weight, offset, w, b, mode, 1.0))
    bm, bb = timeit(lambda: F.spn_backward(gout, init, weight, offset, w, mode, 1.0, need_grad_init=gi))
    fbytes = npx * 29 * es
    bbytes = npx * (28 + 27) * es + (npx * 4 if gi else 0) + npx * es
    print(f"{tag:14s} B={B:5d} {H}x{W} {str(dtype)[6:]:8s} gi={int(gi)} sig={sigma:3.1f} | fwd {fm*1e3:8.1f} us {fbytes/fm/1e6:7.0f} GB/s ({fbytes/fm/1e6/PEAK:5.3f}) "
          f"{npx/fm/1e6:6.2f} Gpix/s | bwd {bm*1e3:8.1f} us {bbytes/bm/1e6:7.0f} GB/s ({bbytes/bm/1e6/PEAK:5.3f})", flush=True)


if __name__ == "__main__":
    f32, bf16 = torch.float32, torch.bfloat16
    for th in ("16", "8", "4"):
        os.environ["JSPSR_SPN_TILE_H"] = th
        run(4096, 128, 128, f32, tag="bench TH=" + th)
    os.environ.pop("JSPSR_SPN_TILE_H")
    run(4096, 128, 128, f32, tag="bench")
    run(4096, 128, 128, f32, gi=True, tag="bench+gi")
    run(4096, 128, 128, bf16, tag="bf16")
    run(4096, 128, 128, f32, mode=2, tag="sum-mode")
    run(4096, 128, 128, f32, sigma=4.0, tag="wide-offsets")
    run(1, 8192, 8192, f32, tag="raster 8k")
    run(1, 4096, 32768, f32, tag="strip 4kx32k")
    run(16, 2000, 2004, f32, tag="generic cs")
    run(64, 334, 334, f32, tag="334 manual")
    for B in (70, 50, 2):
        run(B, 128, 128, f32, tag="config batch")
    # NLSPN affinity front-end (one launch each way)
    B, H, W = 2048, 128, 128
    g = torch.Generator(device="cuda").manual_seed(2)
    conv_out = torch.randn(B, 24, H, W, device="cuda", generator=g)
    conv_out[:, 16:] *= 60
    conf = torch.rand(B, 1, H, W, device="cuda", generator=g)
    gamma = torch.full((1,), 4.0, device="cuda")
    go_ = torch.randn(B, 18, H, W, device="cuda", generator=g); ga_ = torch.randn(B, 9, H, W, device="cuda", generator=g)
    fm, _ = timeit(lambda: F.nlspn_affinity_forward(conv_out, conf, gamma, "TGASS"))
    bm, _ = timeit(lambda: F.nlspn_affinity_backward(go_, ga_, conv_out, conf, gamma, "TGASS"))
    npx = B * H * W
    print(f"nlspn affinity B={B}: fwd {fm*1e3:.1f} us {npx*4*(24+1+27)/fm/1e6:.0f} GB/s ({npx*4*(24+1+27)/fm/1e6/PEAK:.3f}) | "
          f"bwd {bm*1e3:.1f} us {npx*4*(27+24+1+24+1)/bm/1e6:.0f} GB/s ({npx*4*(27+24+1+24+1)/bm/1e6/PEAK:.3f})", flush=True)
    # launch-latency view: both kernels in one CUDA graph (what bench.py reports as config_batch)
    for B in (70, 50, 2):
        g = torch.Generator(device="cuda").manual_seed(1)
        init = torch.rand(B, 1, 128, 128, device="cuda", generator=g)
        weight = torch.rand(B, 9, 128, 128, device="cuda", generator=g)
        offset = 1.5 * torch.randn(B, 18, 128, 128, device="cuda", generator=g)
        gout = torch.randn(B, 1, 128, 128, device="cuda", generator=g)
        w = torch.ones(1, 1, 3, 3, device="cuda"); b = torch.zeros(1, device="cuda")
        s_ = torch.cuda.Stream(); s_.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s_):
            for _ in range(3):
                F.spn_forward(init, weight, offset, w, b, 1, 1.0); F.spn_backward(gout, init, weight, offset, w, 1, 1.0, need_grad_init=False)
        torch.cuda.current_stream().wait_stream(s_)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s_):
            F.spn_forward(init, weight, offset, w, b, 1, 1.0); F.spn_backward(gout, init, weight, offset, w, 1, 1.0, need_grad_init=False)
        m, _ = timeit(gr.replay, n=50, warm=5)
        print(f"graph fwd+bwd B={B}: {m*1e3:.1f} us", flush=True)
