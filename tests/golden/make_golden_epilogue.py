"""Golden vectors for the loss + metric epilogue, produced by the REFERENCE's own classes.

Run in the build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_epilogue.py

* Loss: the reference's `get_loss` + `MultiLoss` (losses/loss_schemes.py:6-33, 55-72) with the YAML weights
  (configs/*.yml:67-70), imported unmodified; autograd gives d Total / d pred.  `kornia` is not installed, so
  `kornia.filters.spatial_gradient` - the one function EdgeLoss takes from it (losses/loss_functions.py:7,182-183) -
  is supplied by `sobel_like_kornia` below, a torch restatement of kornia's published algorithm (replicate pad,
  normalised Sobel pair, [B,C,2,H,W]); everything else on the path is the reference's code.
* Metric: the reference's `MeterRMSE` (evaluation/metrics.py:338-396, package "local") fed one sample at a time as
  the validation loop does (valid_batch_size 1), border 0.05, log de-normalisation with the r8 limits
  (configs/jspsr_r8_img.yml:47-49: min -80, max 929) and linear de-normalisation.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden_tiles import import_reference  # noqa: E402


def sobel_like_kornia(x, mode="sobel", order=1, normalized=True):
    assert mode == "sobel" and order == 1 and normalized
    kx = torch.tensor([[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]], dtype=x.dtype, device=x.device)
    k = torch.stack([kx, kx.t()]) / 8.0
    b, c, h, w = x.shape
    out = F.conv2d(F.pad(x.reshape(b * c, 1, h, w), [1, 1, 1, 1], "replicate"), k[:, None])
    return out.reshape(b, c, 2, h, w)


def main():
    kornia = types.ModuleType("kornia")
    kornia.filters = types.ModuleType("kornia.filters")
    kornia.filters.spatial_gradient = sobel_like_kornia
    sys.modules["kornia"], sys.modules["kornia.filters"] = kornia, kornia.filters
    LS, _ = import_reference("losses.loss_schemes")
    M, stubs = import_reference("evaluation.metrics")
    print("stubbed third-party modules:", sorted(stubs))
    weights = {"L1": 1, "L2": 1, "Grad": 0.1}
    res = {}
    for tag, seed, B, H, W in (("loss_a", 11, 2, 128, 128), ("loss_b", 12, 3, 9, 37), ("loss_c", 13, 1, 1, 5)):
        for dtype, dt in ((torch.float32, "f32"), (torch.float64, "f64")):
            g = torch.Generator().manual_seed(seed)
            gt = torch.rand(B, 1, H, W, generator=g, dtype=torch.float64)
            pred = (gt + 0.05 * torch.randn(B, 1, H, W, generator=g, dtype=torch.float64)).to(dtype).requires_grad_()
            gt = gt.to(dtype)
            crit = LS.MultiLoss(**{k: {"loss_fn": LS.get_loss(k), "weight": v} for k, v in weights.items()})
            out = crit(pred, gt)
            out["Total"].backward()
            if dt == "f32":
                res[f"{tag}_pred"], res[f"{tag}_gt"] = pred.detach().numpy(), gt.numpy()
            for k in ("L1", "L2", "Grad", "Total"):
                res[f"{tag}_{k}_{dt}"] = np.array(out[k].item())
            res[f"{tag}_grad_{dt}"] = pred.grad.numpy()
    meta = [{"subset": "synthetic_x", "id": "a-b-c-d"}]
    for tag, seed, B, H, W, vmin, vmax, elev_log in (("metric_log", 21, 3, 128, 128, -80, 929, True),
                                                     ("metric_lin", 22, 2, 40, 56, -80.0, 929.0, False)):
        g = torch.Generator().manual_seed(seed)
        gt = torch.rand(B, 1, H, W, generator=g)
        pred = gt + 0.03 * torch.randn(B, 1, H, W, generator=g)           # some values leave [0,1]: the clamp matters
        meter = M.MeterRMSE("local", border=0.05, value_min=vmin, value_max=vmax, verbose=False)
        for i in range(B):
            meter.update(pred[i:i + 1], gt[i:i + 1], meta=meta, elev_log=elev_log)
        res[f"{tag}_pred"], res[f"{tag}_gt"] = pred.numpy(), gt.numpy()
        res[f"{tag}_sample_rmse"] = np.array(meter.sample_rmse)
        res[f"{tag}_score"] = np.array(meter.total_rmse / meter.total_n)
        res[f"{tag}_meta"] = np.array([vmin, vmax, float(elev_log), 0.05])
    path = os.path.join(HERE, "epilogue_reference.npz")
    np.savez_compressed(path, **res)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
