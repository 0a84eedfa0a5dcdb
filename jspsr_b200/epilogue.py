"""Loss + metric epilogue of the propagation output over the C ABI in include/jspsr_tiles.h.

* `MultiLoss` - the reference's losses/loss_schemes.py:55-72 for the loss set every YAML config uses
  (configs/*.yml:67-70: L1, L2, Grad): `forward(pred, gt)` returns the same dict {"L1", "L2", "Grad", "Total"};
  `Total` carries the gradient.  One kernel computes the four numbers and dTotal/dpred in a single pass.
* `dem_metrics` / `MeterRMSE` - evaluation/metrics.py:142-199, 338-396 (border crop, clamp, de-normalise, RMSE)
  and the matching MAE, per sample, in one pass.

There is no CPU fallback.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib
from .functional import _count, _ptr, _require_cuda, _stream_ptr, _workspace


def _check_pair(pred: torch.Tensor, gt: torch.Tensor):
    _require_cuda(pred, gt)
    if pred.shape != gt.shape or pred.dim() != 4:
        raise RuntimeError(f"jspsr_b200.epilogue: pred {tuple(pred.shape)} and gt {tuple(gt.shape)} must be equal [B,C,H,W]")
    if pred.dtype != torch.float32 or gt.dtype != torch.float32:
        raise RuntimeError(f"jspsr_b200.epilogue: float32 tensors only, got {pred.dtype} / {gt.dtype}")


def loss_l1_l2_grad(pred, gt, w_l1=1.0, w_l2=1.0, w_grad=0.1, want_grad=True):
    """-> (losses [4] = L1, L2, Grad, Total on device, dTotal/dpred or None)."""
    _check_pair(pred, gt)
    p, g = pred.detach().contiguous(), gt.detach().contiguous()
    B, C, H, W = p.shape
    losses = torch.empty(4, dtype=torch.float32, device=p.device)
    grad = torch.empty_like(p) if want_grad else None
    with torch.cuda.device(p.device):
        _lib.check(_lib.lib().jspsr_loss_l1_l2_grad(_ptr(p), _ptr(g), w_l1, w_l2, w_grad, _ptr(losses), _ptr(grad),
                                                    _ptr(_workspace(p)), B * C, H, W, _stream_ptr(p)),
                   "jspsr_loss_l1_l2_grad")
    _count()
    return losses, grad


class _Loss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, gt, w_l1, w_l2, w_grad):
        need = pred.requires_grad
        losses, grad = loss_l1_l2_grad(pred, gt, w_l1, w_l2, w_grad, want_grad=need)
        ctx.grad = grad
        ctx.mark_non_differentiable(losses)
        return losses[3].clone(), losses

    @staticmethod
    def backward(ctx, g_total, _g_losses):
        # the kernel wrote dTotal/dpred for an upstream gradient of 1; scale by the actual one
        return ctx.grad * g_total, None, None, None, None


class MultiLoss(nn.Module):
    """Drop-in for losses.loss_schemes.MultiLoss (loss_schemes.py:55-83) when the configured losses are the YAML
    configs' {L1, L2, Grad}: `MultiLoss(L1=1, L2=1, Grad=0.1)` or the reference's
    `MultiLoss(**{"L1": {"loss_fn": ..., "weight": 1}, ...})` (the loss_fn entries are ignored)."""

    SUPPORTED = ("L1", "L2", "Grad")

    def __init__(self, **loss_dict):
        super().__init__()
        self.weights = {}
        for name, spec in loss_dict.items():
            key = {"l1": "L1", "l2": "L2", "mse": "L2", "grad": "Grad", "edge": "Grad"}.get(name.lower())
            if key is None:
                raise NotImplementedError(f"jspsr_b200.MultiLoss fuses L1, L2 and Grad only; got {name}")
            self.weights[key] = float(spec["weight"] if isinstance(spec, dict) else spec)
        self.loss_dict = loss_dict
        self.out = {}

    def forward(self, pred, gt):
        w = self.weights
        e = _lib.ext()
        if e is not None:
            _check_pair(pred, gt)
            total, losses = e.multi_loss(pred, gt, w.get("L1", 0.0), w.get("L2", 0.0), w.get("Grad", 0.0))
        else:
            total, losses = _Loss.apply(pred, gt, w.get("L1", 0.0), w.get("L2", 0.0), w.get("Grad", 0.0))
        parts = losses.unbind(0)
        self.out = {k: parts[i] for i, k in enumerate(self.SUPPORTED) if k in w}
        self.out["Total"] = total
        return self.out

    def reset(self):
        self.out = {}


def dem_metrics(pred, gt, border=0.0, value_min=0.0, value_max=1.0, elev_log=False) -> dict:
    """Per-sample {"rmse", "mae", "sum_sq", "sum_abs"} (float64 device tensors [B]) of the de-normalised DEMs."""
    _check_pair(pred, gt)
    if pred.shape[1] != 1:
        raise RuntimeError("jspsr_b200.epilogue: DEM metrics take single-channel tensors [B,1,H,W]")
    p, g = pred.detach().contiguous(), gt.detach().contiguous()
    B, _, H, W = p.shape
    bh, bw = int(H * border), int(W * border)
    sums = torch.empty(B, 2, dtype=torch.float64, device=p.device)
    with torch.cuda.device(p.device):
        _lib.check(_lib.lib().jspsr_dem_metrics(_ptr(p), _ptr(g), _ptr(sums), B, H, W, bh, bw, float(value_min),
                                                float(value_max), int(bool(elev_log)), _stream_ptr(p)),
                   "jspsr_dem_metrics")
    _count(2)  # memset + kernel
    n = (H - 2 * bh) * (W - 2 * bw)
    mean = sums / n  # one launch for both columns (same bits as dividing each column)
    return {"sum_sq": sums[:, 0], "sum_abs": sums[:, 1], "count": n, "rmse": torch.sqrt(mean[:, 0]), "mae": mean[:, 1]}


class MeterRMSE:
    """evaluation/metrics.py:338-421 (package "local") on device tensors: same constructor arguments, `update`,
    `reset`, `get_score`; the per-sample bookkeeping by file name (`meta`) is the caller's."""

    def __init__(self, package="local", tensor_range="[0, 1]", border=0.0, value_min=0.0, value_max=1.0, verbose=True):
        if package != "local" or tensor_range != "[0, 1]":
            raise NotImplementedError
        self.package, self.tensor_range, self.border = package, tensor_range, border
        self.value_min, self.value_max, self.verbose = value_min, value_max, verbose
        self.name = "RMSE"
        self.reset()

    def update(self, pred, gt, meta=None, base_elev=0, elev_log=False):
        # the reference pools the whole batch tensor into ONE rmse per update (metrics.py:382-396: sum / numel, then
        # total_n += 1); its evaluation loop runs at batch size 1, where that is the per-sample rmse
        m = dem_metrics(pred, gt, self.border, self.value_min, self.value_max, elev_log)
        v = float(torch.sqrt(m["sum_sq"].sum() / (m["count"] * pred.shape[0])).item())
        self.total_rmse += v
        self.sample_rmse.append(v)
        self.total_n += 1

    def reset(self):
        self.total_rmse, self.total_n, self.sample_rmse, self.sample_id = 0.0, 0, [], []

    def get_score(self):
        return self.total_rmse / self.total_n
