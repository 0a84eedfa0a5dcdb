"""Golden vectors for the fused Generator tail, produced by the REFERENCE ITSELF.

Run in the build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_generator.py

Builds the reference's own `Generator` (models/components/spn.py:8-75, block = BasicBlock; bc = 32 -> 128 feature
channels is what models/JSPSR.py:28,181-190 builds for the YAML configs' num_feature = 32, bc = 16 -> 64 channels the
cat_only = False / EDSR-with-32-features variant) and `PostProcessor` (spn.py:79-118) with seeded parameters,
runs  weight, offset = generator(dem, context);  out = postprocessor(dem.detach(), weight, offset)  exactly as
models/JSPSR.py:371-375, and records - through a forward hook on `generator.block` - the tensor the fused kernel
starts from (`feature`, spn.py:65), the two 1x1 convolutions' parameters, every intermediate and all gradients,
in fp32 and fp64.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch
import torchvision

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    from models.components.basics import BasicBlock
    from models.components.spn import Generator, PostProcessor
    meta = f"torch {torch.__version__} torchvision {torchvision.__version__}"
    cases = {
        # name: (seed, B, H, W, in_channels, residual, scale, offset gain, bc)
        "gen_residual": (301, 2, 8, 20, 32, True, 1.0, 20.0, 16),
        "gen_sum": (302, 1, 5, 132, 32, False, 1.0, 10.0, 16),
        "gen_residual_half": (303, 1, 9, 128, 32, True, 0.5, 40.0, 16),
        # bc = 32 -> 128 feature channels: what models/JSPSR.py builds at every YAML config (cat_only = True,
        # JSPSR.py:28,181: bc = num_feature = 32)
        "gen_c128_residual": (304, 1, 6, 40, 64, True, 1.0, 20.0, 32),
    }
    only = sys.argv[1:]  # optional: names of the cases to (re)generate
    for name, (seed, B, H, W, cin, residual, scale, gain, bc) in cases.items():
        if only and name not in only:
            continue
        res = {}
        for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            torch.manual_seed(seed)
            with contextlib.redirect_stdout(io.StringIO()):
                gen = Generator(cin, 3, BasicBlock, bc=bc).double()
                pp = PostProcessor(3, residual, scale).double()
            g = torch.Generator().manual_seed(seed)
            with torch.no_grad():
                # an untrained Generator gives |offset| << 1: scale the offset head so taps move
                gen.conv_offset.conv[0].weight.mul_(gain)
                gen.conv_offset.conv[0].bias.copy_(0.5 * torch.randn(16, generator=g, dtype=torch.float64))
                gen.conv_weight[0].weight.mul_(3.0)
                pp.w.copy_(1 + 0.2 * (torch.rand(1, 1, 3, 3, generator=g, dtype=torch.float64) - 0.5))
                pp.b.fill_(0.1)
                for p_ in list(gen.parameters()) + list(pp.parameters()):
                    p_.copy_(p_.float().double())  # fp32-representable parameters for both runs
            dem = torch.rand(B, 1, H, W, generator=g, dtype=torch.float64).float().double()
            context = torch.randn(B, cin, H, W, generator=g, dtype=torch.float64).float().double()
            grad_out = torch.randn(B, 1, H, W, generator=g, dtype=torch.float64).float().double()
            gen, pp = gen.to(dtype), pp.to(dtype)
            dem, context, grad_out = dem.to(dtype), context.to(dtype), grad_out.to(dtype)
            gen.eval()  # BasicBlock holds BatchNorm: freeze it so fp32/fp64 runs see the same statistics path
            captured = {}

            def hook(_m, _inp, outp):
                # the fp64 run continues from the fp32 run's feature values, so both runs (and the kernel)
                # start from the same tensor
                if tag == "f64":
                    outp = torch.from_numpy(res["in_feature"]).double().requires_grad_(True)
                outp.retain_grad()
                captured["feature"] = outp
                return outp

            h = gen.block.register_forward_hook(hook)
            weight, offset = gen(dem, context)
            h.remove()
            weight.retain_grad(); offset.retain_grad()
            out = pp(dem.detach(), weight, offset)
            out.backward(grad_out)
            feature = captured["feature"]
            if tag == "f32":
                res.update(in_init=dem.numpy(), in_grad_out=grad_out.numpy(), in_w=pp.w.detach().numpy(),
                           in_b=pp.b.detach().numpy(),
                           in_conv_weight_w=gen.conv_weight[0].weight.detach().numpy(),
                           in_conv_weight_b=gen.conv_weight[0].bias.detach().numpy(),
                           in_conv_offset_w=gen.conv_offset.conv[0].weight.detach().numpy(),
                           in_conv_offset_b=gen.conv_offset.conv[0].bias.detach().numpy())
            if tag == "f32":
                res["in_feature"] = feature.detach().numpy()   # output of generator.block: the kernel's input
            res[f"{tag}_weight"] = weight.detach().numpy()
            res[f"{tag}_offset"] = offset.detach().numpy()
            res[f"{tag}_out"] = out.detach().numpy()
            res[f"{tag}_grad_feature"] = feature.grad.numpy().astype(np.float32)
            res[f"{tag}_grad_conv_weight_w"] = gen.conv_weight[0].weight.grad.numpy()
            res[f"{tag}_grad_conv_weight_b"] = gen.conv_weight[0].bias.grad.numpy()
            res[f"{tag}_grad_conv_offset_w"] = gen.conv_offset.conv[0].weight.grad.numpy()
            res[f"{tag}_grad_conv_offset_b"] = gen.conv_offset.conv[0].bias.grad.numpy()
            res[f"{tag}_grad_w"] = pp.w.grad.numpy()
            res[f"{tag}_grad_b"] = pp.b.grad.numpy()
        res["norm_mode"] = np.array(1 if residual else 2)
        res["scale"] = np.array(scale)
        res["meta"] = np.array(meta)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **res)
        print("wrote", name, "feature", res["in_feature"].shape, "|offset| max", float(np.abs(res["f64_offset"]).max()),
              "weight range", float(res["f64_weight"].min()), float(res["f64_weight"].max()))


if __name__ == "__main__":
    main()
