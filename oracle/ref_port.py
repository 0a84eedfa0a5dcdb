"""Reference-faithful CPU port of the propagation call sites (TEST / BASELINE
INFRASTRUCTURE ONLY - never imported by jspsr_b200).

The reference's own CPU path is: two torch elementwise ops around the third-party
operator `torchvision.ops.deform_conv2d` (not vendored in the reference; pinned
there to torchvision 0.16, the image ships 0.26).  /root/reference cannot travel
to the GPU box, so bench.py's `cpu_baseline` / `--impl reference` legs time this
restatement of the call sites on the same third-party operator:

  postprocessor_step   <- models/components/spn.py:99-118  (PostProcessor.forward)
                          + `out.backward(grad)` as train/train_utils.py:217 does

tests/test_oracle_golden.py::test_ref_port_matches_fixtures pins it to the fixtures
the real reference produced.
"""
from __future__ import annotations

import torch


def postprocessor_forward(init, weight, offset, w, b, residual=True, scale=1.0):
    from torchvision.ops import deform_conv2d
    if residual:   # spn.py:100-101
        m = weight - torch.mean(weight, 1).unsqueeze(1).expand_as(weight)
    else:          # spn.py:102-103
        m = weight / torch.sum(weight, 1).unsqueeze(1).expand_as(weight)
    out = deform_conv2d(init, offset, weight=w, bias=b, stride=(1, 1), padding=(1, 1), dilation=(1, 1), mask=m)
    if residual:   # spn.py:116-117
        out = out + scale * init
    return out


def postprocessor_step(init, weight, offset, w, b, grad_out, residual=True, scale=1.0):
    """One training step of the hot path on CPU: forward + backward for weight, offset, w, b
    (the DEM is detached in JSPSR, models/JSPSR.py:372)."""
    weight = weight.detach().requires_grad_(True)
    offset = offset.detach().requires_grad_(True)
    w = w.detach().requires_grad_(True)
    b = b.detach().requires_grad_(True)
    out = postprocessor_forward(init.detach(), weight, offset, w, b, residual, scale)
    out.backward(grad_out)
    return out.detach(), weight.grad, offset.grad, w.grad, b.grad
