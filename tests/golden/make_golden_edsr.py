"""EDSR call site: the REFERENCE's own `models.EDSR.EDSR(spn=True)` run end to end on a synthetic DFC30-shaped batch,
with the tensors at the propagation boundary captured from inside it (models/EDSR.py:121-134: `post_layer`, a
PostProcessor(3, True) behind a Generator with bc = n_features // 2 = 16, i.e. C = 64 feature channels).

Run in the build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_edsr.py

* builds `EDSR(in_channels, out_channels=1, n_resblocks=num_block, n_features=num_feature, scale=1, spn=True)` as
  utils/common_config.py:20-38 does (unmodified reference code, seeded default initialisation, train mode, CPU fp32),
  with `num_block: 2, num_feature: 32` of the YAML configs;
* feeds the channel concatenation the EDSR branch receives (lr_dem first: EDSR.py:122-123 takes channel 0 as the DEM);
* hooks `model.post_layer` for (dem, weight, offset) and its output, and `model.generator.block` for the feature the
  fused Generator tail starts from; runs an L1 + L2 loss on the output and backward, and keeps every gradient at that
  boundary (weight, offset, post_layer.w / .b, the feature and the two 1x1 convolutions).
"""
import os
import sys

import numpy as np
import torch
import torchvision

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
from make_golden_tiles import import_reference  # noqa: E402

P = 32   # keeps the committed fixture near 1 MB (the feature and its gradient are 64 channels each)


def main():
    from jspsr_b200 import synth
    E, stubbed = import_reference("models.EDSR")
    torch.manual_seed(515)
    model = E.EDSR(in_channels=4, out_channels=1, n_resblocks=2, n_features=32, scale=1, spn=True)
    model.train()
    with torch.no_grad():   # a trained checkpoint does not have w == 1, b == 0
        model.post_layer.w.add_(0.1 * (torch.rand(1, 1, 3, 3) - 0.5))
        model.post_layer.b.fill_(-0.01)
    batch = synth.dfc30_batch(2, P, resolution=8, with_mask=False, seed=515)
    x = torch.cat([batch["lr_dem"], batch["image"]], dim=1)
    cap = {}

    def pp_hook(mod, inputs, output):
        cap["dem"], cap["weight"], cap["offset"] = inputs
        cap["out"] = output
        for t in (inputs[1], inputs[2], output):
            t.retain_grad()

    def block_hook(mod, inputs, output):
        cap["feature"] = output
        output.retain_grad()

    model.post_layer.register_forward_hook(pp_hook)
    model.generator.block.register_forward_hook(block_hook)
    pred = model(x)
    assert torch.equal(pred, cap["out"]) and not cap["dem"].requires_grad
    loss = (pred - batch["hr_dem"]).abs().mean() + ((pred - batch["hr_dem"]) ** 2).mean()
    loss.backward()
    g = model.generator
    cw, co = g.conv_weight[0], g.conv_offset.conv[0]
    n = lambda t: t.detach().numpy().astype(np.float32)
    arrays = {
        "in_dem": n(cap["dem"]), "in_weight": n(cap["weight"]), "in_offset": n(cap["offset"]),
        "in_w": n(model.post_layer.w), "in_b": n(model.post_layer.b),
        "ref_out": n(cap["out"]), "ref_grad_out": n(cap["out"].grad), "ref_grad_weight": n(cap["weight"].grad),
        "ref_grad_offset": n(cap["offset"].grad), "ref_grad_w": n(model.post_layer.w.grad),
        "ref_grad_b": n(model.post_layer.b.grad),
        "in_feature": n(cap["feature"]), "ref_grad_feature": n(cap["feature"].grad),
        "in_conv_weight_w": n(cw.weight), "in_conv_weight_b": n(cw.bias),
        "in_conv_offset_w": n(co.weight), "in_conv_offset_b": n(co.bias),
        "ref_grad_conv_weight_w": n(cw.weight.grad), "ref_grad_conv_weight_b": n(cw.bias.grad),
        "ref_grad_conv_offset_w": n(co.weight.grad), "ref_grad_conv_offset_b": n(co.bias.grad),
        "residual": np.array(bool(model.post_layer.residual)), "scale": np.array(float(model.post_layer.scale)),
        "meta": np.array(f"torch {torch.__version__} torchvision {torchvision.__version__}"),
    }
    path = os.path.join(HERE, "edsr_spn.npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB; C = {arrays['in_feature'].shape[1]}, "
          f"|offset| std {arrays['in_offset'].std():.2f} max {np.abs(arrays['in_offset']).max():.2f}, "
          f"weight mean {arrays['in_weight'].mean():.2f}, grad_out max {np.abs(arrays['ref_grad_out']).max():.2e}; stubbed {stubbed}")


if __name__ == "__main__":
    main()
