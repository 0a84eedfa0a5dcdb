"""Model-level golden vectors: the REFERENCE's own `models.JSPSR.Model` run end to end on synthetic DFC30-shaped
batches, with the tensors at the propagation boundary captured from inside it.

Run in the build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_model.py

For the four YAML configurations (BASELINE configs 1-4: r3 / r8, image and image+mask guidance) it
* builds `Model(in_channels, num_feature=32, layers=(2,2,2,2), spn=True)` exactly as utils/config.py:50-52,57-60 and
  utils/common_config.py:71-87 do (unmodified reference code, seeded default initialisation), train mode, CPU fp32;
* feeds `jspsr_b200.synth.dfc30_batch(2, 64, resolution, with_mask)` (the synthetic stand-in for data/dfc30.py);
* hooks `model.postprocessor` (models/JSPSR.py:375) for its inputs (dem, weight, offset) and output, and - for
  r8_img - `model.generator.block` (models/components/spn.py:65) for the feature the fused Generator tail starts from;
* runs the configs' loss (losses/loss_schemes.py MultiLoss: L1 + L2 + 0.1 Grad, kornia's Sobel restated as in
  make_golden_epilogue.py) and backward, and keeps every gradient at that boundary:
  d loss/d out, grad_weight, grad_offset, grad of postprocessor.w / .b (and feature / 1x1-conv gradients for r8_img);
* evaluates the reference's `MeterRMSE` (evaluation/metrics.py:361-396, border 0.05, log de-normalisation with the
  config's limits) sample by sample as the validation loop does, plus the MAE of the same de-normalised, cropped,
  clamped tensors (the reference has no MAE meter, SURVEY F10: `mean |pred - gt|` after MeterBase._prepare and
  ToDEM.descale_data).

Patches are 64 x 64 (the U-Net accepts any multiple of 8) instead of the configs' 128 x 128 to keep the committed
fixtures small (about 2 MB per configuration); the channel structure, the network and the arithmetic are the configs'.
"""
import contextlib
import io
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
from make_golden_epilogue import sobel_like_kornia  # noqa: E402

P = 64
CONFIGS = {  # name: (resolution, with_mask, elevation max of the config, seed)
    "r3_img": (3, False, 933.0, 301),          # configs/jspsr_r3_img.yml
    "r8_img": (8, False, 929.0, 801),          # configs/jspsr_r8_img.yml
    "r3_img_msk": (3, True, 933.0, 302),       # configs/jspsr_r3_img_msk.yml
    "r8_img_msk": (8, True, 929.0, 802),       # configs/jspsr_r8_img_msk.yml
}


def main():
    from jspsr_b200 import synth
    kornia = types.ModuleType("kornia")
    kornia.filters = types.ModuleType("kornia.filters")
    kornia.filters.spatial_gradient = sobel_like_kornia
    sys.modules["kornia"], sys.modules["kornia.filters"] = kornia, kornia.filters
    LS, _ = import_reference("losses.loss_schemes")
    M, _ = import_reference("evaluation.metrics")
    DU, _ = import_reference("data.data_utils")
    J, _ = import_reference("models.JSPSR")
    meta = f"torch {torch.__version__} torchvision {torchvision.__version__}"
    torch.set_num_threads(8)

    for name, (res, with_mask, vmax, seed) in CONFIGS.items():
        torch.manual_seed(seed)
        in_channels = {"lr_dem": 1, "COP30": 1, "image": 3}          # utils/config.py:50-52
        if with_mask:
            in_channels["mask"] = 15
        with contextlib.redirect_stdout(io.StringIO()):
            model = J.Model(in_channels=in_channels, num_feature=32, layers=(2, 2, 2, 2), spn=True)
        model.train()
        with torch.no_grad():   # a trained checkpoint does not have w == 1, b == 0
            model.postprocessor.w.add_(0.1 * (torch.rand(1, 1, 3, 3) - 0.5))
            model.postprocessor.b.fill_(0.02)
        batch = synth.dfc30_batch(2, P, resolution=res, with_mask=with_mask, seed=seed)
        cap = {}

        def pp_hook(mod, inputs, output):
            cap["dem"], cap["weight"], cap["offset"] = inputs
            cap["out"] = output
            for t in (inputs[1], inputs[2], output):
                t.retain_grad()

        def block_hook(mod, inputs, output):
            cap["feature"] = output
            output.retain_grad()

        h1 = model.postprocessor.register_forward_hook(pp_hook)
        h2 = model.generator.block.register_forward_hook(block_hook)
        inputs = [batch["lr_dem"], batch["image"]] + ([batch["mask"]] if with_mask else [])
        pred = model(*inputs)
        h1.remove()
        h2.remove()
        weights = {"L1": 1, "L2": 1, "Grad": 0.1}                      # configs/*.yml:67-70
        crit = LS.MultiLoss(**{k: {"loss_fn": LS.get_loss(k), "weight": v} for k, v in weights.items()})
        losses = crit(pred, batch["hr_dem"])
        losses["Total"].backward()

        meter = M.MeterRMSE("local", border=0.05, value_min=synth.ELEV_MIN, value_max=vmax, verbose=False)
        mae = []
        with torch.no_grad():
            for i in range(pred.shape[0]):                             # valid_batch_size 1 (configs/*.yml:96)
                p_i, g_i = pred[i:i + 1].detach(), batch["hr_dem"][i:i + 1]
                meter.update(p_i, g_i, meta=[{"subset": "synthetic_x", "id": "a-b-c-d"}], elev_log=True)
                pc, gc = meter._prepare(p_i, g_i)
                pd_ = DU.ToDEM.descale_data(pc, synth.ELEV_MIN, vmax, True)
                gd_ = DU.ToDEM.descale_data(gc, synth.ELEV_MIN, vmax, True)
                mae.append(float((pd_ - gd_).abs().mean()))

        n = lambda t: t.detach().numpy().astype(np.float32)
        arrays = {
            "in_dem": n(cap["dem"]), "in_weight": n(cap["weight"]), "in_offset": n(cap["offset"]),
            "in_w": n(model.postprocessor.w), "in_b": n(model.postprocessor.b), "in_hr_dem": n(batch["hr_dem"]),
            "ref_out": n(cap["out"]), "ref_grad_out": n(cap["out"].grad), "ref_grad_weight": n(cap["weight"].grad),
            "ref_grad_offset": n(cap["offset"].grad), "ref_grad_w": n(model.postprocessor.w.grad),
            "ref_grad_b": n(model.postprocessor.b.grad),
            "ref_losses": np.array([losses[k].item() for k in ("L1", "L2", "Grad", "Total")]),
            "ref_sample_rmse": np.array(meter.sample_rmse), "ref_rmse": np.array(meter.total_rmse / meter.total_n),
            "ref_sample_mae": np.array(mae), "ref_mae": np.array(float(np.mean(mae))),
            "cfg": np.array([res, float(with_mask), synth.ELEV_MIN, vmax, 0.05]), "meta": np.array(meta),
            "residual": np.array(bool(model.postprocessor.residual)), "scale": np.array(float(model.postprocessor.scale)),
        }
        assert torch.equal(cap["out"], pred)
        if name == "r8_img":     # the Generator-tail level of the same run (jspsr_b200.generator_postprocess's boundary)
            g = model.generator
            cw, co = g.conv_weight[0], g.conv_offset.conv[0]
            arrays.update({
                "in_feature": n(cap["feature"]), "ref_grad_feature": n(cap["feature"].grad),
                "in_conv_weight_w": n(cw.weight), "in_conv_weight_b": n(cw.bias),
                "in_conv_offset_w": n(co.weight), "in_conv_offset_b": n(co.bias),
                "ref_grad_conv_weight_w": n(cw.weight.grad), "ref_grad_conv_weight_b": n(cw.bias.grad),
                "ref_grad_conv_offset_w": n(co.weight.grad), "ref_grad_conv_offset_b": n(co.bias.grad),
            })
        path = os.path.join(HERE, f"model_{name}.npz")
        np.savez_compressed(path, **arrays)
        print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB; losses {arrays['ref_losses']}, "
              f"rmse {arrays['ref_sample_rmse']}, mae {arrays['ref_sample_mae']}, "
              f"|offset| max {np.abs(arrays['in_offset']).max():.2f}, grad_out max {np.abs(arrays['ref_grad_out']).max():.2e}")


if __name__ == "__main__":
    main()
