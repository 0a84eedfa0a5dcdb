import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jspsr_b200 import functional as F
g = torch.Generator(device="cuda").manual_seed(1)
B, H, W = 6, 128, 128
init = torch.rand(B, 1, H, W, device="cuda", generator=g)
weight = torch.sigmoid(1.5 * torch.randn(B, 9, H, W, device="cuda", generator=g))
offset = (1.5 * torch.randn(B, 18, H, W, device="cuda", generator=g)).clamp_(-8, 8)
gout = torch.randn(B, 1, H, W, device="cuda", generator=g)
w = 1 + 0.1 * torch.randn(1, 1, 3, 3, device="cuda"); b = torch.full((1,), 0.1, device="cuda")
res = {}
for th in ("16", "8", "4", "2"):
    os.environ["JSPSR_SPN_TILE_H"] = th
    o = F.spn_forward(init, weight, offset, w, b, 1, 1.0)
    gr = F.spn_backward(gout, init, weight, offset, w, 1, 1.0, need_grad_init=True)
    res[th] = (o, gr)
for th in ("8", "4", "2"):
    o, gr = res[th]; o0, gr0 = res["16"]
    d = (o - o0).abs()
    print("TH", th, "fwd equal", torch.equal(o, o0), "max diff", d.max().item(), "n diff", int((d > 0).sum()),
          "| gw equal", torch.equal(gr[1], gr0[1]), "go equal", torch.equal(gr[2], gr0[2]),
          "gi maxdiff", (gr[0] - gr0[0]).abs().max().item())
    if not torch.equal(o, o0):
        idx = (d > 0).nonzero()[:5]
        print("  first diffs at", idx.tolist(), [ (o[tuple(i)].item(), o0[tuple(i)].item()) for i in idx])
