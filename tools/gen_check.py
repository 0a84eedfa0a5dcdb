"""Dev tool: fused generator tail vs torch (GPU fp32, TF32 off / CPU fp64) + timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jspsr_b200 import functional as F
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

def ref(init, feat, cw, cb, w, b, mode, scale, dt=torch.float64):
    init, feat, cw, cb, w, b = (t.detach().cpu().to(dt) for t in (init, feat, cw, cb, w, b))
    z = torch.einsum("nc,bchw->bnhw", cw, feat) + cb.view(1, -1, 1, 1)
    weight = torch.sigmoid(z[:, :9])
    B, _, H, W = init.shape
    o = z[:, 9:]
    offset = torch.cat((o[:, :8], torch.zeros(B, 2, H, W, dtype=dt), o[:, 8:]), 1)
    return weight, offset

def main():
    torch.manual_seed(0)
    for (B, H, W) in ((2, 128, 128), (1, 40, 200), (3, 17, 64)):
        C = 64
        init = torch.rand(B, 1, H, W, device="cuda")
        feat = torch.randn(B, C, H, W, device="cuda")
        cw = torch.randn(25, C, device="cuda") * 0.15
        cb = torch.randn(25, device="cuda") * 0.1
        w = torch.ones(1, 1, 3, 3, device="cuda") + 0.05 * torch.randn(1, 1, 3, 3, device="cuda"); b = torch.full((1,), 0.1, device="cuda")
        out, weight, offset = F.gen_spn_forward(init, feat, cw, cb, w, b, 1, 1.0, True)
        out2 = F.gen_spn_forward(init, feat, cw, cb, w, b, 1, 1.0, False)
        torch.cuda.synchronize()
        rw, ro = ref(init, feat, cw, cb, w, b, 1, 1.0)
        print(f"B={B} {H}x{W}: weight err {(weight.cpu().double()-rw).abs().max():.3e}  offset err {(offset.cpu().double()-ro).abs().max():.3e} "
              f"(|offset| max {ro.abs().max():.2f})  out==out2 {torch.equal(out, out2)}")
        out_ref = F.spn_forward(init, rw.float().cuda(), ro.float().cuda(), w, b, 1, 1.0)
        out_same = F.spn_forward(init, weight, offset, w, b, 1, 1.0)
        print(f"    out vs unfused(exact w/o) {(out-out_ref).abs().max():.3e}   out vs unfused(kernel's own w/o) {(out-out_same).abs().max():.3e}")
    # timing
    from tools.quick_bench import timeit, PEAK
    B, H, W, C = 2048, 128, 128, 64
    init = torch.rand(B, 1, H, W, device="cuda"); feat = torch.randn(B, C, H, W, device="cuda")
    cw = torch.randn(25, C, device="cuda") * 0.15; cb = torch.randn(25, device="cuda") * 0.1
    cw[9:] *= 1.2
    w = torch.ones(1, 1, 3, 3, device="cuda"); b = torch.zeros(1, device="cuda")
    npx = B * H * W
    for th in ("16", "8"):
        os.environ["JSPSR_SPN_TILE_H"] = th
        for halo in ("narrow", "wide"):
            os.environ["JSPSR_SPN_HALO"] = halo
            try:
                m, _ = timeit(lambda: F.gen_spn_forward(init, feat, cw, cb, w, b, 1, 1.0, False))
                m2, _ = timeit(lambda: F.gen_spn_forward(init, feat, cw, cb, w, b, 1, 1.0, True))
                by, by2 = npx * (C * 4 + 8), npx * (C * 4 + 8 + 108)
                print(f"TH={th} {halo}: fused fwd {m*1e3:.1f} us {by/m/1e6:.0f} GB/s ({by/m/1e6/PEAK:.3f}) | +weight/offset out {m2*1e3:.1f} us {by2/m2/1e6:.0f} GB/s ({by2/m2/1e6/PEAK:.3f})", flush=True)
            except Exception as e:
                print("TH", th, halo, "failed:", e)
    os.environ.pop("JSPSR_SPN_TILE_H"); os.environ.pop("JSPSR_SPN_HALO")
    # unfused incumbent: torch 1x1 convs (cuDNN/cuBLAS, TF32 allowed as torch's default) + sigmoid + cat + our forward
    torch.backends.cudnn.allow_tf32 = True
    cwt = cw[:9].reshape(9, C, 1, 1).contiguous(); cot = cw[9:].reshape(16, C, 1, 1).contiguous()
    def unfused():
        weight = torch.sigmoid(torch.nn.functional.conv2d(feat, cwt, cb[:9]))
        o = torch.nn.functional.conv2d(feat, cot, cb[9:])
        o = o.view(B, 8, 2, H, W)
        lo = list(torch.chunk(o, 8, dim=1)); lo.insert(4, torch.zeros((B, 1, 2, H, W), device="cuda"))
        offset = torch.cat(lo, dim=1).view(B, -1, H, W)
        return F.spn_forward(init, weight, offset, w, b, 1, 1.0)
    m, _ = timeit(unfused)
    print(f"unfused (torch convs, TF32 allowed, + our forward): {m*1e3:.1f} us")

main()
