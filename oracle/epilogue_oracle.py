"""CPU oracle (numpy) for the loss + metric epilogue.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
leg as the checker; the product (jspsr_b200/) never imports it.

Restates, with the reference lines it follows:

* multi_loss - losses/loss_schemes.py:55-72 (MultiLoss: every configured loss on (pred, gt), `Total` = sum of
  weight * loss) with the YAML configs' losses (configs/*.yml:67-70: L1 1, L2 1, Grad 0.1): L1 = nn.L1Loss,
  L2 = nn.MSELoss (loss_schemes.py:8-11), Grad = EdgeLoss (losses/loss_functions.py:171-185):
  L1Loss(spatial_gradient(pred), spatial_gradient(gt)).
* spatial_gradient - `kornia.filters.spatial_gradient` (mode "sobel", order 1, normalized) is a THIRD-PARTY
  dependency that is absent from /root/reference and from this image (the reference pins no version; its container
  nvcr.io/nvidia/pytorch:23.10-py3 + `pip install kornia`, ReadMe.md:9-19).  Its published algorithm: replicate-pad
  by one pixel, correlate with the Sobel pair [[-1,0,1],[-2,0,2],[-1,0,1]] (d/dx) and its transpose (d/dy), each
  divided by the sum of absolute values (8), output [B,C,2,H,W].  PARITY UNPINNED for this one function: no copy of
  kornia exists here to run; the fixtures pin everything around it (the reference's own MultiLoss / L1Loss / MSELoss
  / EdgeLoss classes run on a torch restatement of the same published algorithm, tests/golden/make_golden_epilogue.py).
* multi_loss_grad - the analytic gradient of `Total` w.r.t. pred (what autograd computes for the reference).
* dem_metrics - evaluation/metrics.py:142-199 (MeterBase._prepare: crop int(h*border), clamp pred to [0,1]) and
  :361-382 (MeterRMSE.update: ToDEM.descale_data on both, sqrt(sum(d^2)/numel)), data/data_utils.py:441-457
  (descale_data: exp(x*log(max-min))+min in log mode), plus the matching mean absolute error.
"""
from __future__ import annotations

from math import log

import numpy as np

SOBEL_X = np.array([[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]]) / 8.0
SOBEL_Y = SOBEL_X.T.copy()


def spatial_gradient(x: np.ndarray) -> np.ndarray:
    """x [B,C,H,W] -> [B,C,2,H,W] (dx, dy), replicate border, normalised Sobel."""
    B, C, H, W = x.shape
    p = np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1)), mode="edge")
    out = np.zeros((B, C, 2, H, W), x.dtype)
    for i in range(3):
        for j in range(3):
            win = p[:, :, i:i + H, j:j + W]
            out[:, :, 0] += x.dtype.type(SOBEL_X[i, j]) * win
            out[:, :, 1] += x.dtype.type(SOBEL_Y[i, j]) * win
    return out


def multi_loss(pred: np.ndarray, gt: np.ndarray, w_l1=1.0, w_l2=1.0, w_grad=0.1) -> dict:
    d = pred - gt
    out = {"L1": np.abs(d).mean(), "L2": (d * d).mean(),
           "Grad": np.abs(spatial_gradient(pred) - spatial_gradient(gt)).mean()}
    out["Total"] = w_l1 * out["L1"] + w_l2 * out["L2"] + w_grad * out["Grad"]
    return out


def multi_loss_grad(pred: np.ndarray, gt: np.ndarray, w_l1=1.0, w_l2=1.0, w_grad=0.1) -> np.ndarray:
    """d Total / d pred."""
    B, C, H, W = pred.shape
    n = pred.size
    d = pred - gt
    g = (w_l1 * np.sign(d) + w_l2 * 2.0 * d) / n
    s = np.sign(spatial_gradient(pred) - spatial_gradient(gt)) * (w_grad / (2.0 * n))     # [B,C,2,H,W]
    gp = np.zeros((B, C, H + 2, W + 2), pred.dtype)                                      # gradient of the padded image
    for i in range(3):
        for j in range(3):
            gp[:, :, i:i + H, j:j + W] += SOBEL_X[i, j] * s[:, :, 0] + SOBEL_Y[i, j] * s[:, :, 1]
    # replicate padding folds the border ring back onto the edge pixels
    gp[:, :, 1, :] += gp[:, :, 0, :]
    gp[:, :, H, :] += gp[:, :, H + 1, :]
    gp[:, :, :, 1] += gp[:, :, :, 0]
    gp[:, :, :, W] += gp[:, :, :, W + 1]
    return g + gp[:, :, 1:H + 1, 1:W + 1]


def descale(data, elev_min, elev_max, elev_log=False):
    if elev_log:
        return np.exp(data * data.dtype.type(log(elev_max - elev_min))) + data.dtype.type(elev_min)
    return data * data.dtype.type(elev_max - elev_min) + data.dtype.type(elev_min)


def dem_metrics(pred, gt, border=0.05, value_min=0.0, value_max=1.0, elev_log=False):
    """Per-sample (sum d^2, sum |d|, count, rmse, mae) of the de-normalised, border-cropped, clamped DEMs."""
    h, w = pred.shape[-2:]
    bh, bw = int(h * border), int(w * border)
    p = np.clip(pred[..., bh:h - bh, bw:w - bw], 0.0, 1.0)
    t = gt[..., bh:h - bh, bw:w - bw]
    d = (descale(p, value_min, value_max, elev_log) - descale(t, value_min, value_max, elev_log))
    d = d.reshape(d.shape[0], -1).astype(np.float64)
    n = d.shape[1]
    sq, ab = (d * d).sum(axis=1), np.abs(d).sum(axis=1)
    return {"sum_sq": sq, "sum_abs": ab, "count": n, "rmse": np.sqrt(sq / n), "mae": ab / n}
