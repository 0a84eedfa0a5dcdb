"""Model-level parity (BASELINE configs 1-4): tensors captured INSIDE the reference's own `models.JSPSR.Model` run
(tests/golden/make_golden_model.py: r3 / r8, image and image+mask guidance, synthetic DFC30-shaped batch, the configs'
loss and backward, the reference's MeterRMSE) replayed through the CUDA path:

* `jspsr_b200.PostProcessor` on the captured (dem, weight, offset): out and, driven by the reference's own d loss/d out,
  grad_weight / grad_offset / grad_w / grad_b within 1e-5 of each tensor's own scale (the upstream gradient of a
  mean-reduced loss is ~1e-4 here, so the old `max(1, scale)` rule would have been vacuous);
* the chain the training loop runs (train/train_utils.py:211-217): PostProcessor -> MultiLoss(L1 + L2 + 0.1 Grad):
  the four loss values within 1e-5 relative;
* evaluation/metrics.py:361-396 on the model output: RMSE per sample and the score, and the MAE of the same
  de-normalised tensors, IDENTICAL at the 4 decimals the reference prints;
* for r8_img the same run one level up: `functional.gen_propagate` from the captured Generator feature
  (models/components/spn.py:65) with the reference's 1x1-convolution parameters: out, grad_feature and the
  convolution gradients (models/JSPSR.py:371-375).
"""
import glob
import os

import numpy as np
import pytest
import torch

from tests.test_gpu_parity import ABS_FLOOR, FP32_TOL, assert_close, dev  # noqa: F401  (same gate)

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODEL = sorted(glob.glob(os.path.join(GOLDEN, "model_*.npz")))


@pytest.fixture(scope="module")
def jb():
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    import jspsr_b200
    from jspsr_b200 import _lib
    _lib.lib()
    return jspsr_b200


def test_fixtures_present():
    assert [os.path.basename(p) for p in MODEL] == ["model_r3_img.npz", "model_r3_img_msk.npz", "model_r8_img.npz",
                                                    "model_r8_img_msk.npz"]


def _postprocessor(jb, z):
    pp = jb.PostProcessor(3, bool(z["residual"]), float(z["scale"])).cuda()
    with torch.no_grad():
        pp.w.copy_(dev(z["in_w"]))
        pp.b.copy_(dev(z["in_b"]))
    return pp


@pytest.mark.parametrize("path", MODEL, ids=[os.path.basename(p)[6:-4] for p in MODEL])
def test_postprocessor_inside_the_reference_model(jb, path):
    z = np.load(path)
    pp = _postprocessor(jb, z)
    dem, hr = dev(z["in_dem"]), dev(z["in_hr_dem"])
    weight, offset = dev(z["in_weight"]).requires_grad_(), dev(z["in_offset"]).requires_grad_()
    out = pp(dem, weight, offset)
    assert_close(out, z["ref_out"], FP32_TOL, "out vs the reference Model's output")
    # backward driven by the reference's own d loss / d out: isolates the propagation (the loss's sign() kinks are
    # decided by rounding and are tested in test_gpu_tiles_epilogue.py)
    gout = z["ref_grad_out"]
    out.backward(dev(gout))
    assert_close(weight.grad, z["ref_grad_weight"], FP32_TOL, "grad_weight", gout=gout)
    assert_close(offset.grad, z["ref_grad_offset"], FP32_TOL, "grad_offset", gout=gout)
    assert_close(pp.w.grad, z["ref_grad_w"], FP32_TOL, "grad_w", gout=gout)
    assert_close(pp.b.grad, z["ref_grad_b"], FP32_TOL, "grad_b", gout=gout)

    # the training chain: propagation -> the configs' loss
    crit = jb.MultiLoss(L1=1.0, L2=1.0, Grad=0.1)
    with torch.no_grad():
        losses = crit(pp(dem, weight.detach(), offset.detach()), hr)
    got = np.array([float(losses[k]) for k in ("L1", "L2", "Grad", "Total")])
    assert np.all(np.abs(got - z["ref_losses"]) <= 1e-5 * np.abs(z["ref_losses"])), (got, z["ref_losses"])

    # evaluation: RMSE (the reference's meter, one sample per update) and MAE, identical at the printed 4 decimals
    _, _, vmin, vmax, border = (float(v) for v in z["cfg"])
    meter = jb.MeterRMSE("local", border=border, value_min=vmin, value_max=vmax, verbose=False)
    pred = out.detach()
    for i in range(pred.shape[0]):
        meter.update(pred[i:i + 1], hr[i:i + 1], elev_log=True)
    m = jb.epilogue.dem_metrics(pred, hr, border, vmin, vmax, True)
    for i in range(pred.shape[0]):
        assert f"{meter.sample_rmse[i]:.4f}" == f"{float(z['ref_sample_rmse'][i]):.4f}"
        assert f"{float(m['mae'][i]):.4f}" == f"{float(z['ref_sample_mae'][i]):.4f}"
    assert f"{meter.get_score():.4f}" == f"{float(z['ref_rmse']):.4f}"
    assert f"{float(m['mae'].mean()):.4f}" == f"{float(z['ref_mae']):.4f}"


def test_generator_tail_inside_the_reference_model(jb):
    """r8_img one level up: the fused Generator tail + propagation from the feature captured at generator.block."""
    from jspsr_b200 import functional as F
    z = np.load(os.path.join(GOLDEN, "model_r8_img.npz"))
    C = z["in_feature"].shape[1]
    assert C == 128   # models/JSPSR.py:28,181: cat_only -> bc = num_feature = 32 -> 4 * bc feature channels
    feature = dev(z["in_feature"]).requires_grad_()
    conv_w = dev(np.concatenate([z["in_conv_weight_w"].reshape(9, C), z["in_conv_offset_w"].reshape(16, C)])).requires_grad_()
    conv_b = dev(np.concatenate([z["in_conv_weight_b"], z["in_conv_offset_b"]])).requires_grad_()
    w, b = dev(z["in_w"]).requires_grad_(), dev(z["in_b"]).requires_grad_()
    out = F.gen_propagate(dev(z["in_dem"]), feature, conv_w, conv_b, w, b, 1 if bool(z["residual"]) else 2, float(z["scale"]))
    assert_close(out, z["ref_out"], FP32_TOL, "fused tail: out vs the reference Model's output")
    gout = z["ref_grad_out"]
    out.backward(dev(gout))
    ref_cw = np.concatenate([z["ref_grad_conv_weight_w"].reshape(9, C), z["ref_grad_conv_offset_w"].reshape(16, C)])
    ref_cb = np.concatenate([z["ref_grad_conv_weight_b"], z["ref_grad_conv_offset_b"]])
    # the reference's gradients went through its fp32 convolutions: twice the tolerance, like the gen_* fixtures
    assert_close(feature.grad, z["ref_grad_feature"], 2 * FP32_TOL, "grad_feature", gout=gout)
    assert_close(conv_w.grad, ref_cw, 2 * FP32_TOL, "grad_conv_w", gout=gout)
    assert_close(conv_b.grad, ref_cb, 2 * FP32_TOL, "grad_conv_b", gout=gout)
    assert_close(w.grad, z["ref_grad_w"], 2 * FP32_TOL, "grad_w", gout=gout)
    assert_close(b.grad, z["ref_grad_b"], 2 * FP32_TOL, "grad_b", gout=gout)


def test_edsr_call_site_inside_the_reference_model(jb):
    """tests/golden/edsr_spn.npz (make_golden_edsr.py): models/EDSR.py:121-134 captured inside the reference's own
    EDSR(spn=True) run - `post_layer` = PostProcessor(3, True) behind a Generator with bc = 16, i.e. the fused tail at
    C = 64 feature channels (two CTAs per SM, the other tcgen05 instantiation than the JSPSR configs' C = 128)."""
    from jspsr_b200 import functional as F
    z = np.load(os.path.join(GOLDEN, "edsr_spn.npz"))
    gout = z["ref_grad_out"]
    pp = _postprocessor(jb, z)
    weight, offset = dev(z["in_weight"]).requires_grad_(), dev(z["in_offset"]).requires_grad_()
    out = pp(dev(z["in_dem"]), weight, offset)
    assert_close(out, z["ref_out"], FP32_TOL, "post_layer: out vs the reference EDSR's output")
    out.backward(dev(gout))
    assert_close(weight.grad, z["ref_grad_weight"], FP32_TOL, "grad_weight", gout=gout)
    assert_close(offset.grad, z["ref_grad_offset"], FP32_TOL, "grad_offset", gout=gout)
    assert_close(pp.w.grad, z["ref_grad_w"], FP32_TOL, "grad_w", gout=gout)
    assert_close(pp.b.grad, z["ref_grad_b"], FP32_TOL, "grad_b", gout=gout)

    C = z["in_feature"].shape[1]
    assert C == 64
    feature = dev(z["in_feature"]).requires_grad_()
    conv_w = dev(np.concatenate([z["in_conv_weight_w"].reshape(9, C), z["in_conv_offset_w"].reshape(16, C)])).requires_grad_()
    conv_b = dev(np.concatenate([z["in_conv_weight_b"], z["in_conv_offset_b"]])).requires_grad_()
    w, b = dev(z["in_w"]).requires_grad_(), dev(z["in_b"]).requires_grad_()
    fused = F.gen_propagate(dev(z["in_dem"]), feature, conv_w, conv_b, w, b, 1, 1.0)
    assert_close(fused, z["ref_out"], FP32_TOL, "fused tail (C = 64): out vs the reference EDSR's output")
    fused.backward(dev(gout))
    ref_cw = np.concatenate([z["ref_grad_conv_weight_w"].reshape(9, C), z["ref_grad_conv_offset_w"].reshape(16, C)])
    ref_cb = np.concatenate([z["ref_grad_conv_weight_b"], z["ref_grad_conv_offset_b"]])
    assert_close(feature.grad, z["ref_grad_feature"], 2 * FP32_TOL, "grad_feature", gout=gout)
    assert_close(conv_w.grad, ref_cw, 2 * FP32_TOL, "grad_conv_w", gout=gout)
    assert_close(conv_b.grad, ref_cb, 2 * FP32_TOL, "grad_conv_b", gout=gout)
    assert_close(w.grad, z["ref_grad_w"], 2 * FP32_TOL, "grad_w", gout=gout)
    assert_close(b.grad, z["ref_grad_b"], 2 * FP32_TOL, "grad_b", gout=gout)
