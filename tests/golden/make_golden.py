"""Generate golden input/output vectors by running the REFERENCE ITSELF on CPU.

Run in the build container only (needs /root/reference, which does not exist
on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

It imports the reference's own modules
  models.components.spn.PostProcessor      (spn.py:79-118)
  models.LRRU.Post_process_deconv          (LRRU.py:250-298)
  models.components.nlspn.NLSPN            (nlspn.py:8-235)
unmodified from /root/reference, feeds them seeded inputs in fp32 and fp64,
and stores inputs, outputs and all gradients as small .npz fixtures next to
this script.  The fixtures pin oracle/spn_oracle.py, oracle/spn_oracle.c and
(through them, and directly) the CUDA path.
"""
import os
import sys
import types

import numpy as np
import torch
import torchvision

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    from models.components.spn import PostProcessor
    from models.components.nlspn import NLSPN
    from models.LRRU import Post_process_deconv
    return PostProcessor, Post_process_deconv, NLSPN


def _inputs(seed, B, H, W, off_sigma, dtype, mode="normal", gout_scale=1.0):
    g = torch.Generator().manual_seed(seed)
    init = torch.rand(B, 1, H, W, generator=g, dtype=torch.float64)
    weight = torch.sigmoid(1.5 * torch.randn(B, 9, H, W, generator=g, dtype=torch.float64))
    offset = off_sigma * torch.randn(B, 18, H, W, generator=g, dtype=torch.float64)
    if mode == "zero":
        offset.zero_()
    elif mode == "integer":
        offset = torch.round(offset)
    elif mode == "half":
        offset = torch.round(offset * 2) / 2
    offset[:, 8:10] = 0          # Generator inserts a zero centre pair (spn.py:69-73)
    grad_out = gout_scale * torch.randn(B, 1, H, W, generator=g, dtype=torch.float64)
    w = 1 + 0.2 * (torch.rand(1, 1, 3, 3, generator=g, dtype=torch.float64) - 0.5)
    b = torch.tensor([0.1], dtype=torch.float64)
    # inputs are rounded to fp32 first so fp32 and fp64 runs see identical values
    cast = lambda t: t.to(torch.float32).to(dtype)
    return [cast(t) for t in (init, weight, offset, grad_out, w, b)]


def run_postprocessor(make_module, seed, B, H, W, off_sigma, mode, dtype, gout_scale=1.0):
    init, weight, offset, grad_out, w, b = _inputs(seed, B, H, W, off_sigma, dtype, mode, gout_scale)
    mod = make_module().to(dtype)
    with torch.no_grad():
        mod.w.copy_(w)
        mod.b.copy_(b)
    init.requires_grad_(True)
    weight.requires_grad_(True)
    offset.requires_grad_(True)
    out = mod(init, weight, offset)
    out.backward(grad_out)
    return dict(out=out.detach(), grad_init=init.grad, grad_weight=weight.grad,
                grad_offset=offset.grad, grad_w=mod.w.grad, grad_b=mod.b.grad), \
        dict(init=init.detach(), weight=weight.detach(), offset=offset.detach(),
             grad_out=grad_out, w=w, b=b)


def main():
    PostProcessor, Post_process_deconv, NLSPN = _import_reference()
    meta = f"torch {torch.__version__} torchvision {torchvision.__version__}"
    print(meta)

    import io, contextlib
    pp_cases = {
        # name: (factory, seed, B, H, W, sigma, offset-mode, norm_mode, scale)
        "pp_residual": (lambda: PostProcessor(3, True, 1.0), 11, 2, 12, 16, 1.5, "normal", 1, 1.0),
        "pp_residual_scale": (lambda: PostProcessor(3, True, 0.5), 12, 2, 9, 13, 1.5, "normal", 1, 0.5),
        "pp_sum": (lambda: PostProcessor(3, False, 1.0), 13, 2, 12, 16, 1.5, "normal", 2, 1.0),
        "pp_far_offsets": (lambda: PostProcessor(3, True, 1.0), 14, 2, 10, 12, 8.0, "normal", 1, 1.0),
        "pp_zero_offsets": (lambda: PostProcessor(3, True, 1.0), 15, 1, 8, 8, 0.0, "zero", 1, 1.0),
        "pp_integer_offsets": (lambda: PostProcessor(3, False, 1.0), 16, 1, 9, 11, 2.0, "integer", 2, 1.0),
        "pp_half_offsets": (lambda: PostProcessor(3, True, 1.0), 17, 1, 9, 11, 2.0, "half", 1, 1.0),
        "pp_w32_multirow": (lambda: PostProcessor(3, True, 1.0), 18, 1, 40, 32, 2.5, "normal", 1, 1.0),
        # the gradient a mean-reduced loss hands back (train/train_utils.py:214-217: losses are means over ~1e6 pixels):
        # grad_out ~ 1e-6, so every gradient tensor is ~1e-6 and only a tolerance relative to the tensor's own scale tests it
        "pp_small_grad": (lambda: PostProcessor(3, True, 1.0), 19, 2, 24, 32, 1.5, "normal", 1, 1.0, 1e-6),
        "pp_small_grad_sum": (lambda: PostProcessor(3, False, 1.0), 20, 2, 24, 32, 1.5, "normal", 2, 1.0, 1e-6),
        "lrru_residual": (lambda: Post_process_deconv(types.SimpleNamespace(kernel_size=3, dkn_residual=True)),
                          21, 2, 12, 16, 1.5, "normal", 1, 1.0),
        "lrru_sum": (lambda: Post_process_deconv(types.SimpleNamespace(kernel_size=3, dkn_residual=False)),
                     22, 2, 12, 16, 1.5, "normal", 2, 1.0),
    }
    for name, (factory, seed, B, H, W, sigma, omode, nmode, scale, *rest) in pp_cases.items():
        gs = rest[0] if rest else 1.0

        def quiet_factory():
            with contextlib.redirect_stdout(io.StringIO()):
                return factory()
        out32, inp = run_postprocessor(quiet_factory, seed, B, H, W, sigma, omode, torch.float32, gs)
        out64, _ = run_postprocessor(quiet_factory, seed, B, H, W, sigma, omode, torch.float64, gs)
        arrays = {f"in_{k}": v.numpy() for k, v in inp.items()}
        arrays.update({f"f32_{k}": v.numpy() for k, v in out32.items()})
        arrays.update({f"f64_{k}": v.numpy() for k, v in out64.items()})
        arrays["norm_mode"] = np.array(nmode)
        arrays["scale"] = np.array(scale)
        arrays["meta"] = np.array(meta)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
        print("wrote", name, {k: v.shape for k, v in arrays.items() if k.startswith("f32_")})

    # ---- NLSPN ----------------------------------------------------------
    nl_cases = {
        "nlspn_tgass_conf": dict(affinity="TGASS", conf_prop=True, preserve_input=False, legacy=False, T=3),
        "nlspn_tgass_noconf": dict(affinity="TGASS", conf_prop=False, preserve_input=False, legacy=False, T=6),
        "nlspn_as": dict(affinity="AS", conf_prop=True, preserve_input=False, legacy=False, T=2),
        "nlspn_ass_preserve": dict(affinity="ASS", conf_prop=True, preserve_input=True, legacy=False, T=2),
        "nlspn_tc_legacy": dict(affinity="TC", conf_prop=True, preserve_input=False, legacy=True, T=2),
    }
    for i, (name, cfg) in enumerate(nl_cases.items()):
        res = {}
        for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            g = torch.Generator().manual_seed(100 + i)
            B, H, W, ch_g = 2, 10, 14, 8
            args = types.SimpleNamespace(prop_time=cfg["T"], affinity=cfg["affinity"], affinity_gamma=0.5,
                                         conf_prop=cfg["conf_prop"], preserve_input=cfg["preserve_input"],
                                         legacy=cfg["legacy"])
            mod = NLSPN(args, ch_g, 1, 3, 3)
            cw = 0.3 * torch.randn(mod.conv_offset_aff.weight.shape, generator=g, dtype=torch.float64)
            cb = 0.5 * torch.randn(mod.conv_offset_aff.bias.shape, generator=g, dtype=torch.float64)
            # aff channels are divided by 100 inside tanh: scale them up so tanh is exercised
            cw[16:] *= 60
            cb[16:] *= 60
            guidance = torch.randn(B, ch_g, H, W, generator=g, dtype=torch.float64)
            confidence = torch.rand(B, 1, H, W, generator=g, dtype=torch.float64)
            feat_init = torch.rand(B, 1, H, W, generator=g, dtype=torch.float64)
            feat_fix = torch.rand(B, 1, H, W, generator=g, dtype=torch.float64)
            feat_fix = feat_fix * (torch.rand(B, 1, H, W, generator=g, dtype=torch.float64) > 0.7)
            grad_out = torch.randn(B, 1, H, W, generator=g, dtype=torch.float64)
            grad_mid = torch.randn(B, 1, H, W, generator=g, dtype=torch.float64)
            cast = lambda t: t.to(torch.float32).to(dtype)
            cw, cb, guidance, confidence, feat_init, feat_fix, grad_out, grad_mid = map(
                cast, (cw, cb, guidance, confidence, feat_init, feat_fix, grad_out, grad_mid))
            mod = mod.to(dtype)
            with torch.no_grad():
                mod.conv_offset_aff.weight.copy_(cw)
                mod.conv_offset_aff.bias.copy_(cb)
            # legacy mode shifts the (detached, storage-sharing) offsets IN PLACE
            # (nlspn.py:118-128), which autograd rejects: it is an inference-only
            # switch in the reference, so that fixture is forward-only.
            with_grad = not cfg["legacy"]
            guidance.requires_grad_(with_grad)
            confidence.requires_grad_(with_grad)
            feat_init.requires_grad_(with_grad)
            conv_out = mod.conv_offset_aff(guidance).detach()
            with torch.set_grad_enabled(with_grad):
                feat, list_feat, offset, aff, gamma = mod(feat_init, guidance, confidence, feat_fix, None)
            if with_grad:
                # loss touches the last and the first intermediate so that the T-loop
                # backward with several live outputs is pinned too
                loss = (feat * grad_out).sum() + (list_feat[0] * grad_mid).sum()
                loss.backward()
            else:
                for t in (feat_init, guidance, confidence, mod.conv_offset_aff.weight, mod.conv_offset_aff.bias):
                    t.grad = torch.zeros_like(t)
            if tag == "f32":
                res.update(in_conv_w=cw.numpy(), in_conv_b=cb.numpy(), in_guidance=guidance.detach().numpy(),
                           in_confidence=confidence.detach().numpy(), in_feat_init=feat_init.detach().numpy(),
                           in_feat_fix=feat_fix.numpy(), in_grad_out=grad_out.numpy(), in_grad_mid=grad_mid.numpy(),
                           in_gamma=mod.aff_scale_const.detach().numpy())
            res[f"{tag}_conv_out"] = conv_out.numpy()
            res[f"{tag}_feat"] = feat.detach().numpy()
            res[f"{tag}_list_feat"] = torch.stack(list_feat).detach().numpy()
            res[f"{tag}_offset"] = offset.detach().numpy()
            res[f"{tag}_aff"] = aff.detach().numpy()
            res[f"{tag}_grad_feat_init"] = feat_init.grad.numpy()
            res[f"{tag}_grad_guidance"] = guidance.grad.numpy()
            res[f"{tag}_grad_confidence"] = (confidence.grad if confidence.grad is not None
                                             else torch.zeros_like(confidence)).numpy()
            res[f"{tag}_grad_conv_w"] = mod.conv_offset_aff.weight.grad.numpy()
            res[f"{tag}_grad_conv_b"] = mod.conv_offset_aff.bias.grad.numpy()
            res[f"{tag}_grad_gamma"] = (mod.aff_scale_const.grad if mod.aff_scale_const.grad is not None
                                        else torch.zeros(1, dtype=dtype)).numpy()
        for k, v in cfg.items():
            res[f"cfg_{k}"] = np.array(v)
        res["meta"] = np.array(meta)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **res)
        print("wrote", name)


if __name__ == "__main__":
    main()
