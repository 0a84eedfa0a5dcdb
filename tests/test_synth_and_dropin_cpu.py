"""CPU checks of the synthetic DFC30-shaped generator and, when the reference tree is present (build container
only), of the interface-level drop-in: the replacement module can be installed into the reference's own
models.JSPSR.Model and keeps its state_dict / optimizer-group contract.  No kernel runs here."""
import os
import sys

import numpy as np
import pytest
import torch

from jspsr_b200 import synth

REF = "/root/reference"


def test_dfc30_batch_shapes_and_ranges():
    b = synth.dfc30_batch(4, 128, resolution=8, with_mask=True, seed=3)
    assert tuple(b["lr_dem"].shape) == (4, 1, 128, 128) and tuple(b["hr_dem"].shape) == (4, 1, 128, 128)
    assert tuple(b["image"].shape) == (4, 3, 128, 128) and tuple(b["mask"].shape) == (4, 15, 128, 128)
    for k in ("lr_dem", "hr_dem", "image", "mask"):
        assert b[k].dtype == torch.float32 and float(b[k].min()) >= 0.0 and float(b[k].max()) <= 1.0
    for i in range(15):  # channel i takes the values {0, (i+1)/16}  (data_utils.py:262-265)
        vals = torch.unique(b["mask"][:, i])
        assert set(np.round(vals.numpy(), 6)) <= {0.0, round((i + 1) / 16, 6)}
    assert len(b["meta"]) == 4 and {"id", "subset", "base", "shape", "bbox", "augmentation"} <= set(b["meta"][0])
    # the low-resolution input is a smoothed copy of the target, not noise
    err = (synth.descale_elevation(b["lr_dem"]) - synth.descale_elevation(b["hr_dem"])).abs().mean()
    assert 0.05 < float(err) < 20.0
    again = synth.dfc30_batch(4, 128, resolution=8, with_mask=True, seed=3)
    assert torch.equal(again["hr_dem"], b["hr_dem"])
    assert "mask" not in synth.dfc30_batch(2, 64, resolution=3, seed=1)


def test_scale_descale_round_trip_matches_oracle_arithmetic():
    from oracle import spn_oracle as O
    x = torch.rand(2, 1, 16, 16, generator=torch.Generator().manual_seed(0)) * 0.9 + 0.05
    m = synth.descale_elevation(x, 8).numpy()
    np.testing.assert_allclose(m, O.descale(x.numpy(), synth.ELEV_MIN, synth.ELEV_MAX[8], elev_log=True), rtol=1e-4)
    back = synth.scale_elevation(torch.from_numpy(m), 8)
    np.testing.assert_allclose(back.numpy(), x.numpy(), atol=1e-4)  # fp32 exp/log round trip over a 1009 m range


def test_propagation_inputs_statistics():
    init, weight, offset, gout = synth.propagation_inputs(8, 64, 64, device="cpu")
    assert tuple(offset.shape) == (8, 18, 64, 64) and torch.all(offset[:, 8:10] == 0)
    assert 0.35 < float(weight.mean()) < 0.65 and float(offset.abs().max()) <= 8.0
    assert 1.3 < float(offset[:, :8].std()) < 1.7


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference tree only exists in the build container")
def test_installs_into_the_reference_model():
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import contextlib
    import io
    from models.JSPSR import Model
    import jspsr_b200 as jb
    in_channels = {"lr_dem": 1, "COP30": 1, "image": 3}   # utils/config.py:50-52
    with contextlib.redirect_stdout(io.StringIO()):
        model = Model(in_channels=in_channels, num_feature=32, layers=(2, 2, 2, 2), spn=True)
    ref_pp = model.postprocessor
    ref_keys = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    model.postprocessor = jb.PostProcessor(kernel_size=3, residual=ref_pp.residual, scale=ref_pp.scale)
    new_keys = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert new_keys == ref_keys, "state_dict keys/shapes changed: released checkpoints would not load"
    assert (model.postprocessor.stride, model.postprocessor.padding, model.postprocessor.dilation) == \
           (ref_pp.stride, ref_pp.padding, ref_pp.dilation)
    # a checkpoint written by the reference module loads into the replacement
    sd = {"w": torch.full((1, 1, 3, 3), 0.7), "b": torch.full((1,), -0.2)}
    ref_pp.load_state_dict(sd)
    model.postprocessor.load_state_dict(ref_pp.state_dict())
    assert torch.equal(model.postprocessor.w, ref_pp.w) and torch.equal(model.postprocessor.b, ref_pp.b)
    # the diff_lr optimizer groups pick the layer up by name (utils/common_config.py:250-253)
    names = [n for n, _ in model.named_parameters() if "postprocessor" in n]
    assert names == ["postprocessor.w", "postprocessor.b"]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference tree only exists in the build container")
def test_installs_into_the_reference_edsr_and_lrru_models():
    """The other two construction sites: models/EDSR.py:107 `self.post_layer = PostProcessor(3, True)` and
    models/LRRU.py:399 `self.Post_process = Post_process_deconv(args)` - the replacement modules leave the models'
    state_dict (keys, shapes) as the reference wrote it, so its checkpoints load."""
    import types
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from models.EDSR import EDSR
    from models.LRRU import Model as LRRU
    import jspsr_b200 as jb
    edsr = EDSR(in_channels=4, out_channels=1, n_resblocks=2, n_features=32, scale=1, spn=True)   # common_config.py:20-38
    ref_keys = {k: tuple(v.shape) for k, v in edsr.state_dict().items()}
    ref_pp = edsr.post_layer
    edsr.post_layer = jb.PostProcessor(3, True)
    assert {k: tuple(v.shape) for k, v in edsr.state_dict().items()} == ref_keys
    assert (edsr.post_layer.residual, edsr.post_layer.scale) == (ref_pp.residual, ref_pp.scale)
    edsr.post_layer.load_state_dict(ref_pp.state_dict())
    assert [n for n, _ in edsr.named_parameters() if n.startswith("post_layer")] == ["post_layer.w", "post_layer.b"]

    args = types.SimpleNamespace(input_channels={"lr_dem": 1, "image": 3}, output_channels=1, kernel_size=3, bc=4,
                                 prob=1.0, dkn_residual=True)                                       # common_config.py:57-68
    lrru = LRRU(args)
    ref_keys = {k: tuple(v.shape) for k, v in lrru.state_dict().items()}
    ref_post = lrru.Post_process
    lrru.Post_process = jb.Post_process_deconv(args)
    assert {k: tuple(v.shape) for k, v in lrru.state_dict().items()} == ref_keys
    assert lrru.Post_process.dkn_residual == ref_post.dkn_residual
    lrru.Post_process.load_state_dict(ref_post.state_dict())
    assert [n for n, _ in lrru.named_parameters() if n.startswith("Post_process")] == ["Post_process.w", "Post_process.b"]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference tree only exists in the build container")
def test_generator_postprocess_matches_the_reference_generators_structure(monkeypatch):
    """jspsr_b200.generator_postprocess drives the reference's OWN Generator (models/components/spn.py:8-75): every
    sub-module it touches must exist there with the shapes the fused kernel expects, its body up to `block` must be
    the reference's forward (same feature tensor), and the operand matrix it hands to the kernel must be the two 1x1
    convolutions' parameters in the documented order.  The kernel call is intercepted: no GPU here."""
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import contextlib
    import io
    from models.JSPSR import Model
    import jspsr_b200 as jb
    from jspsr_b200 import functional as F
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = Model(in_channels={"lr_dem": 1, "COP30": 1, "image": 3}, num_feature=32, layers=(2, 2, 2, 2), spn=True)
    gen, pp = model.generator.eval(), model.postprocessor
    # models/JSPSR.py:28 hard-codes cat_only = True, so bc = num_feature = 32 (JSPSR.py:181) and the Generator's feature
    # has bc * 4 = 128 channels at every YAML config (num_feature: 32): the kernel's C = 128 instantiation
    assert gen.kernel_size == 3 and gen.conv_weight[0].weight.shape == (9, 128, 1, 1)
    assert gen.conv_offset.conv[0].weight.shape == (16, 128, 1, 1) and gen.conv_offset.conv[0].bias is not None
    captured = {}
    hook = gen.block.register_forward_hook(lambda _m, _i, o: captured.__setitem__("feature", o.detach()))
    dem, ctx = torch.rand(1, 1, 16, 24), torch.randn(1, gen.convf1.conv[0].in_channels, 16, 24)
    with torch.no_grad():
        weight, offset = gen(dem, ctx)                                   # the reference's own forward
    hook.remove()
    seen = {}

    def fake_kernel(init, feature, conv_w, conv_b, w, b, mode, scale):
        seen.update(init=init, feature=feature, conv_w=conv_w, conv_b=conv_b, mode=mode, scale=scale)
        return init
    monkeypatch.setattr(F, "gen_propagate", fake_kernel)
    with torch.no_grad():
        jb.generator_postprocess(gen, pp, dem, ctx)
    assert torch.equal(seen["feature"], captured["feature"]) and torch.equal(seen["init"], dem)
    assert seen["mode"] == 1 and seen["scale"] == float(pp.scale)
    # the operand reproduces the reference's weight/offset from that feature (spn.py:66-73)
    z = torch.einsum("nc,bchw->bnhw", seen["conv_w"], seen["feature"]) + seen["conv_b"].view(1, -1, 1, 1)
    assert torch.allclose(torch.sigmoid(z[:, :9]), weight, atol=1e-6)
    assert torch.allclose(z[:, 9:17], offset[:, :8], atol=1e-5) and torch.allclose(z[:, 17:], offset[:, 10:], atol=1e-5)
    assert torch.all(offset[:, 8:10] == 0)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference tree only exists in the build container")
def test_generator_postprocess_drives_the_reference_lrru_encoder(monkeypatch):
    """The LRRU twin: the reference's OWN models.LRRU.BasicDepthEncoder (LRRU.py:202-247) + Post_process_deconv through
    jspsr_b200.generator_postprocess - `ref` instead of `block`, plain nn.Conv2d heads, functional sigmoid, bc = 16 (64
    feature channels), `dkn_residual` instead of `residual`, no scale.  Kernel call intercepted: no GPU here."""
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import types
    from models.LRRU import BasicDepthEncoder, Post_process_deconv as RefPost
    import jspsr_b200 as jb
    from jspsr_b200 import functional as F
    torch.manual_seed(1)
    enc = BasicDepthEncoder(kernel_size=3, bc=16, norm_layer=torch.nn.BatchNorm2d).eval()
    args = types.SimpleNamespace(kernel_size=3, dkn_residual=True)
    ref_pp, pp = RefPost(args), jb.Post_process_deconv(args)
    assert set(ref_pp.state_dict()) == set(pp.state_dict())
    captured = {}
    hook = enc.ref.register_forward_hook(lambda _m, _i, o: captured.__setitem__("feature", o.detach()))
    depth, ctx = torch.rand(1, 1, 16, 24), torch.randn(1, 32, 16, 24)
    with torch.no_grad():
        weight, offset = enc(depth, ctx)                                 # the reference's own forward
    hook.remove()
    seen = {}

    def fake_kernel(init, feature, conv_w, conv_b, w, b, mode, scale):
        seen.update(init=init, feature=feature, conv_w=conv_w, conv_b=conv_b, mode=mode, scale=scale)
        return init
    monkeypatch.setattr(F, "gen_propagate", fake_kernel)
    with torch.no_grad():
        jb.generator_postprocess(enc, pp, depth, ctx)
    assert torch.equal(seen["feature"], captured["feature"]) and seen["feature"].shape[1] == 64
    assert seen["mode"] == 1 and seen["scale"] == 1.0
    z = torch.einsum("nc,bchw->bnhw", seen["conv_w"], seen["feature"]) + seen["conv_b"].view(1, -1, 1, 1)
    assert torch.allclose(torch.sigmoid(z[:, :9]), weight, atol=1e-6)
    assert torch.allclose(z[:, 9:17], offset[:, :8], atol=1e-5) and torch.allclose(z[:, 17:], offset[:, 10:], atol=1e-5)
