"""Small run through every kernel family for compute-sanitizer (memcheck / racecheck)."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jspsr_b200
from jspsr_b200 import functional as F
torch.manual_seed(0)
for dt in (torch.float32, torch.bfloat16):
    for (B, H, W, sig) in ((2, 40, 128, 1.5), (1, 37, 150, 6.0), (1, 20, 260, 2.0), (3, 9, 7, 1.0)):
        init = torch.rand(B, 1, H, W, device="cuda").to(dt).requires_grad_()
        weight = torch.rand(B, 9, H, W, device="cuda").to(dt).requires_grad_()
        offset = (sig * torch.randn(B, 18, H, W, device="cuda")).to(dt).requires_grad_()
        for mode_mod in (jspsr_b200.PostProcessor(3, True, 0.5), jspsr_b200.PostProcessor(3, False, 1.0)):
            m = mode_mod.cuda()
            out = m(init, weight, offset)
            out.float().square().mean().backward()
        for tma in ("0", "1"):
            os.environ["JSPSR_SPN_DISABLE_TMA"] = tma
            F.spn_forward(init.detach(), weight.detach(), offset.detach(), m.w, m.b, 1, 1.0)
        os.environ["JSPSR_SPN_DISABLE_TMA"] = "0"
        for halo in ("narrow", "wide"):
            os.environ["JSPSR_SPN_HALO"] = halo
            F.spn_backward(torch.randn_like(init), init.detach(), weight.detach(), offset.detach(), m.w, 1, 1.0)
        os.environ.pop("JSPSR_SPN_HALO")
        F.spn_iterate(init.detach(), weight.detach() * 0.1, offset.detach(), 3)
        F.offset_absmax(offset.detach())
args = types.SimpleNamespace(prop_time=3, affinity="TGASS", affinity_gamma=0.5, conf_prop=True, preserve_input=True, legacy=False)
nl = jspsr_b200.NLSPN(args, 8, 1, 3, 3).cuda()
with torch.no_grad():
    nl.conv_offset_aff.weight.normal_(0, 0.3); nl.conv_offset_aff.bias.normal_(0, 0.5)
g = torch.randn(2, 8, 33, 140, device="cuda", requires_grad=True)
c = torch.rand(2, 1, 33, 140, device="cuda", requires_grad=True)
f0 = torch.rand(2, 1, 33, 140, device="cuda", requires_grad=True)
fix = torch.rand(2, 1, 33, 140, device="cuda") * (torch.rand(2, 1, 33, 140, device="cuda") > 0.7)
feat, lst, off, aff, gam = nl(f0, g, c, fix)
(feat.sum() + lst[0].sum()).backward()
# strips
ti = torch.rand(1, 1, 64, 128, device="cuda"); tw = torch.rand(1, 9, 64, 128, device="cuda"); to = torch.randn(1, 18, 64, 128, device="cuda")
st = torch.zeros(1, dtype=torch.int32, device="cuda")
F.spn_forward_strip(ti[:, :, 10:50].contiguous(), tw[:, :, 16:40].contiguous(), to[:, :, 16:40].contiguous(), None, None, 0, 0.0, 64, 16, 10, st)
# mixed dtype (torch.autocast), generator-tail kernels (tcgen05), feature gradient
for (B, H, W) in ((2, 40, 128), (1, 21, 200), (1, 5, 7)):
    init = torch.rand(B, 1, H, W, device="cuda")
    wb = torch.rand(B, 9, H, W, device="cuda").bfloat16().requires_grad_()
    ob = torch.randn(B, 18, H, W, device="cuda").bfloat16().requires_grad_()
    pp = jspsr_b200.PostProcessor(3, True, 1.0).cuda()
    pp(init, wb, ob).square().mean().backward()
    for C in (64, 128):
        for fdt in (torch.float32, torch.bfloat16):
            feat = torch.randn(B, C, H, W, device="cuda").to(fdt).requires_grad_()
            cw = (0.15 * torch.randn(25, C, device="cuda")).requires_grad_(); cb = (0.1 * torch.randn(25, device="cuda")).requires_grad_()
            out = F.gen_propagate(init, feat, cw, cb, pp.w, pp.b, 1, 1.0)
            out.square().mean().backward()
            F.gen_spn_forward(init, feat.detach(), cw.detach(), cb.detach(), pp.w, pp.b, 2, 1.0)
# backward of the fixed-affinity loop, split form (iter_carry_kernel / iter_grad_kernel): TMA and manual staging,
# offsets that leave the narrow staged tile, T = 1 / 3 / 8, with and without the gradient of the feature
from jspsr_b200 import epilogue as EP
for (B, H, W, T, sig) in ((2, 40, 128, 3, 1.5), (1, 37, 150, 8, 6.0), (1, 128, 128, 1, 1.5), (3, 9, 7, 2, 1.0)):
    fa = torch.rand(B, 1, H, W, device="cuda").requires_grad_()
    aa = (0.2 * torch.randn(B, 9, H, W, device="cuda")).requires_grad_()
    oa = (sig * torch.randn(B, 18, H, W, device="cuda")).requires_grad_()
    F.iterate(fa, aa, oa, T).square().sum().backward()
    F.iterate(fa.detach(), aa, oa, T)[-1].sum().backward()
    # the LRRU blend and the RMSE / MAE sums (float4 and scalar kernels; border columns that are not 16-byte aligned)
    d = fa.detach() * (torch.rand_like(fa) > 0.5)
    F.preserve_blend(fa.detach(), d)
    F.preserve_blend(fa.detach().bfloat16(), d.bfloat16())
    for border in (0.0, 0.05, 0.2):
        if int(H * border) * 2 < H and int(W * border) * 2 < W:
            EP.dem_metrics(fa.detach(), d, border, -80.0, 929.0, True)
            EP.dem_metrics(fa.detach(), d, border, -80.0, 929.0, False)
torch.cuda.synchronize()
print("sanitize case done")
