"""Synthetic DFC30-shaped batches (there is no network access to the real DFC30 data).

Shapes, value ranges and the `meta` record follow what the reference's dataset + transforms deliver to the
model (data/dfc30.py:194-257 `__getitem__`, :347-364 `collate_fn`; data/data_utils.py:262-265 mask scaling,
:289-312 `ToTensor.scale_data`): per batch
    lr_dem, hr_dem  [B,1,P,P] float32 in [0,1]  (log min-max of relative elevation, configs/*.yml:45-52)
    image           [B,3,P,P] float32 in [0,1]
    mask            [B,15,P,P] float32, channel i in {0, (i+1)/16}      (image+mask configs only)
    meta            list of dicts with id / subset / base / shape / bbox / augmentation
The DEMs are band-limited fractal terrain so that the low-resolution input really is a smoothed copy of the target.
`propagation_inputs` draws (init, weight, offset) with the statistics the untrained Generator produces
(weight = sigmoid(N(0,1.5^2)), offsets N(0,1.5^2) clipped to +-8, zero centre pair).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch

ELEV_MIN = -80.0  # configs/*.yml tensor_kwargs.min
ELEV_MAX = {3: 933.0, 8: 929.0}  # tensor_kwargs.max per resolution


def _fractal(B: int, P: int, gen: torch.Generator, device, beta: float = 2.2) -> torch.Tensor:
    """[B,P,P] terrain with a power-law spectrum, zero mean, unit-ish variance."""
    fy = torch.fft.fftfreq(P, device=device)[:, None]
    fx = torch.fft.rfftfreq(P, device=device)[None, :]
    amp = (fx * fx + fy * fy).clamp_min(1.0 / (P * P)) ** (-beta / 2)
    amp[0, 0] = 0
    re = torch.randn(B, P, P // 2 + 1, generator=gen, device=device)
    im = torch.randn(B, P, P // 2 + 1, generator=gen, device=device)
    z = torch.fft.irfft2(torch.complex(re, im) * amp, s=(P, P))
    return z / z.flatten(1).std(dim=1)[:, None, None].clamp_min(1e-6)


def scale_elevation(elev: torch.Tensor, resolution: int = 8) -> torch.Tensor:
    """ToTensor.scale_data with elev_log=True (data_utils.py:289-312)."""
    return torch.log(elev - ELEV_MIN) / math.log(ELEV_MAX[resolution] - ELEV_MIN) + 1e-8


def descale_elevation(x: torch.Tensor, resolution: int = 8) -> torch.Tensor:
    """ToDEM.descale_data with elev_log=True (data_utils.py:441-457)."""
    return torch.exp(x * math.log(ELEV_MAX[resolution] - ELEV_MIN)) + ELEV_MIN


def dfc30_batch(B: int, patch: int = 128, resolution: int = 8, with_mask: bool = False, seed: int = 0,
                device="cpu") -> Dict[str, object]:
    gen = torch.Generator(device=device).manual_seed(seed)
    relief = 60.0 * _fractal(B, patch, gen, device)                    # metres, relative to the tile minimum
    hr = relief - relief.flatten(1).min(dim=1).values[:, None, None] + 1.0   # >= 1 m above the base (assert in scale_data)
    k = resolution if resolution < patch else 1                            # 30 m COP30 cell vs 3 / 8 m target
    lr = torch.nn.functional.avg_pool2d(hr[:, None], k, k)
    lr = torch.nn.functional.interpolate(lr, size=(patch, patch), mode="bicubic", align_corners=False)[:, 0]
    lr = lr.clamp_min(1.0) + 0.5 * torch.randn(B, patch, patch, generator=gen, device=device).clamp(-2, 2)
    lr = lr.clamp_min(1.0)
    batch = {
        "lr_dem": scale_elevation(lr[:, None], resolution).float().clamp(0, 1),
        "hr_dem": scale_elevation(hr[:, None], resolution).float().clamp(0, 1),
        "image": torch.rand(B, 3, patch, patch, generator=gen, device=device),
    }
    if with_mask:
        m = (torch.rand(B, 15, patch, patch, generator=gen, device=device) > 0.8).float()
        scale = (torch.arange(15, device=device, dtype=torch.float32) + 1) / 16.0   # data_utils.py:262-265
        batch["mask"] = m * scale[None, :, None, None]
    meta: List[dict] = []
    for i in range(B):
        meta.append({"id": f"synthetic-dfc30-{seed:04d}-{i:04d}", "subset": "synthetic_train",
                     "shape": (patch, patch, 1 + 3 + (15 if with_mask else 0)),
                     "augmentation": {"rot90": 0, "flip_lr": False, "flip_ud": False},
                     "bbox": (0, 0, patch, patch), "base": float(relief[i].min()), "profile": None})
    batch["meta"] = meta
    return batch


def propagation_inputs(B: int, H: int = 128, W: int = 128, seed: int = 1234, device="cuda", dtype=torch.float32,
                       offset_sigma: float = 1.5, init: Optional[torch.Tensor] = None):
    """(init, weight, offset, grad_out) with the statistics of an untrained Generator."""
    g = torch.Generator(device=device).manual_seed(seed)
    if init is None:
        init = torch.rand(B, 1, H, W, device=device, generator=g)
    weight = torch.sigmoid(1.5 * torch.randn(B, 9, H, W, device=device, generator=g))
    offset = (offset_sigma * torch.randn(B, 18, H, W, device=device, generator=g)).clamp_(-8, 8)
    offset[:, 8:10] = 0
    grad_out = torch.randn(B, 1, H, W, device=device, generator=g)
    return [t.to(dtype) for t in (init, weight, offset, grad_out)]
