"""LRRU cascade golden vectors: the REFERENCE's own `models.LRRU.Model` run on CPU, with every tensor at the
propagation boundary of its four stages captured from inside its forward (models/LRRU.py:447-498).

Run in the build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_lrru.py

* builds `Model(args)` unmodified (input_channels lr_dem + image, kernel_size 3, dkn_residual True, bc = 4 to keep the
  network small; seeded default initialisation, eval mode so that the stochastic-depth blocks are deterministic);
  the last 1x1 convolution of each `weight_offset{i}` (BasicDepthEncoder.conv_offset) is scaled up so that the
  untrained encoders emit offsets of a few pixels (parameter values, not code, are changed);
* feeds a sparse depth map (about 40 % valid pixels, zeros elsewhere) and an image, 2 x 32 x 32;
* hooks `Post_process` (LRRU.py:455, 469, 483, 498) for its inputs (x_i, weight_i, offset_i) and output, and the
  first argument of `weight_offset{i}` - the tensor AFTER the input-preservation blend and `.detach()` of
  LRRU.py:447-453, 460-466, 474-480, 488-494.
So `blend_i` is what the reference makes of (`out_{i-1}`, `d_clear`), and `out_i` what it makes of
(`blend_i`, `weight_i`, `offset_i`): the fixtures of `jspsr_preserve_blend` and of the cascade as a whole.
"""
import os
import sys
import types

import numpy as np
import torch
import torchvision

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
from make_golden_tiles import import_reference  # noqa: E402


def main():
    L, stubbed = import_reference("models.LRRU")
    torch.manual_seed(4711)
    args = types.SimpleNamespace(input_channels={"lr_dem": 1, "image": 3}, output_channels=1, kernel_size=3, bc=4,
                                 prob=0.5, dkn_residual=True)
    model = L.Model(args).eval()
    with torch.no_grad():
        for i in range(4):
            enc = getattr(model, f"weight_offset{i}")
            for m in enc.conv_offset.modules():
                if isinstance(m, torch.nn.Conv2d):
                    m.weight.mul_((5.0, 4.0, 2.0, 1.0)[i])   # -> offsets with a standard deviation of about 2 pixels
    cap = {"pp": []}

    def pre(name):
        def hook(mod, inp):
            cap[name] = inp[0].detach().clone()
        return hook
    for i in range(4):
        getattr(model, f"weight_offset{i}").register_forward_pre_hook(pre(f"blend{i}"))
    model.Post_process.register_forward_hook(
        lambda mod, inp, out: cap["pp"].append(([t.detach().clone() for t in inp], out.detach().clone())))
    g = torch.Generator().manual_seed(4712)
    B, H, W = 2, 32, 32
    dep = torch.rand(B, 1, H, W, generator=g) * (torch.rand(B, 1, H, W, generator=g) > 0.6)
    img = torch.rand(B, 3, H, W, generator=g)
    with torch.no_grad():
        final = model(dep, img)
    assert len(cap["pp"]) == 4
    out = {"d_clear": dep.numpy(), "final": final.numpy(),
           "meta": np.array(f"torch {torch.__version__} torchvision {torchvision.__version__}"),
           "stubbed": np.array(",".join(stubbed))}
    for i, ((x, w, o), y) in enumerate(cap["pp"]):
        assert torch.equal(x, cap[f"blend{i}"])       # Post_process consumes exactly the blended, detached tensor
        out[f"blend{i}"] = x.numpy()
        out[f"weight{i}"] = w.numpy()
        out[f"offset{i}"] = o.numpy()
        out[f"out{i}"] = y.numpy()
        print(i, "offset std", float(o.std()), "max", float(o.abs().max()), "weight mean", float(w.mean()))
    assert torch.equal(final, cap["pp"][3][1])
    path = os.path.join(HERE, "cascade_lrru.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KB; stubbed:", stubbed)


if __name__ == "__main__":
    main()
