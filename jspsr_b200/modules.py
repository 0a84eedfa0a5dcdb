"""Drop-in nn.Modules for the reference's propagation layers.

Same constructor arguments, attributes, forward signatures, return values and
state_dict keys as
  models/components/spn.py:79-118     PostProcessor
  models/LRRU.py:250-298              Post_process_deconv
  models/components/nlspn.py:8-235    NLSPN
so `models/JSPSR.py:192-194,375`, `models/EDSR.py:107,134`,
`models/LRRU.py:399,455-498` and `models/CompletionFormer.py:34-36,59-61` can use
them unchanged and released checkpoints load (`utils/utils.py:360-364` filters by
key name and shape: `w` [1,1,k,k], `b` [1]).  The arithmetic runs in
libjspsr_spn.so; these classes hold parameters and call it.
"""
from __future__ import annotations

from abc import ABC

import torch
import torch.nn as nn

from . import functional as F
from ._lib import NORM_RESIDUAL, NORM_SUM


def _require_k3(kernel_size: int):
    if kernel_size != 3:
        raise NotImplementedError(
            f"jspsr_b200 implements the 3x3 propagation window the reference uses everywhere "
            f"(got kernel_size={kernel_size})")


class PostProcessor(nn.Module):
    """models/components/spn.py:79-118."""

    def __init__(self, kernel_size=3, residual=True, scale=1.0):
        super().__init__()
        _require_k3(kernel_size)
        self.residual = residual
        self.w = nn.Parameter(torch.ones((1, 1, kernel_size, kernel_size)))
        self.b = nn.Parameter(torch.zeros(1))
        self.stride = (1, 1)
        self.padding = ((kernel_size - 1) // 2, (kernel_size - 1) // 2)
        self.dilation = (1, 1)
        self.scale = scale
        if self.scale != 1:
            print("Warning: The scale factor is not 1. This may lead to unexpected results.")

    def set_grad_reducer(self, reducer):
        """Batch-sharded training (one process per GPU): all-reduce the gradients of `w` / `b` inside the backward kernel
        over `reducer`'s ranks (jspsr_b200.peer.PeerGradReducer) instead of through a NCCL bucket.  None switches it off.
        Not a parameter or buffer: state_dict and checkpoints are unchanged."""
        object.__setattr__(self, "_grad_reducer", reducer)
        return self

    def forward(self, init_dem, weight, offset):
        mode = NORM_RESIDUAL if self.residual else NORM_SUM
        return F.propagate(init_dem, weight, offset, self.w, self.b, mode, float(self.scale),
                           getattr(self, "_grad_reducer", None))


def generator_postprocess(generator: nn.Module, postprocessor: nn.Module, dem, context, init_dem=None,
                          detach_dem: bool = True):
    """The two lines `weight, offset = self.generator(dem, context)` and
    `out = self.postprocessor(dem.detach(), weight, offset)` of models/JSPSR.py:371-375 (models/EDSR.py:133-134)
    with the Generator's last two layers fused into the propagation kernel.

    `generator` is the reference's own models.components.spn.Generator (or any module with the same
    sub-modules: convd1, convd2, convf1, convf2, conv, block, conv_weight, conv_offset); its body up to `block`
    (spn.py:57-65: cuDNN convolutions, not on the hot path) runs as is, its parameters stay where they are, so
    checkpoints and optimizer groups are unchanged.  `postprocessor` is a PostProcessor (either implementation).
    The twin pair of the LRRU baseline is accepted as well: models/LRRU.py:202-247 `BasicDepthEncoder` (last body block
    `ref` instead of `block`, plain nn.Conv2d heads with a functional sigmoid, bc = 16 -> 64 feature channels) with
    models/LRRU.py:250-298 `Post_process_deconv` (`dkn_residual`, no scale), as used four times per forward at
    LRRU.py:454-498 (`weight_i, offset_i = self.weight_offset_i(x, ctx); x = self.Post_process(x, weight_i, offset_i)`).
    The weight/offset tensors never exist in inference; with autograd they are written once by the fused kernel
    and consumed by the fused backward.

    `detach_dem`:
      True  (default) what both reference call sites do - models/JSPSR.py:372 `dem = dem.detach()` BEFORE the
            Generator, models/EDSR.py:122-123 `x_copy = x.clone().detach(); dem = x_copy[:, 0:1]`: no gradient reaches
            `dem`, neither through the Generator's convd1 nor through the propagation;
      False for a caller whose `dem` carries a gradient (the LRRU-style cascade without its `.detach()`, NLSPN-style
            refinement of a predicted DEM): it then receives the gradient through the Generator body AND the
            propagation's grad_init.
    `init_dem` overrides the DEM the propagation starts from (used as given, with its own autograd history)."""
    if detach_dem:
        dem = dem.detach()
    d2 = generator.convd2(generator.convd1(dem))
    f2 = generator.convf2(generator.convf1(context))
    last = generator.block if hasattr(generator, "block") else generator.ref           # spn.Generator / LRRU twin
    feature = last(generator.conv(torch.cat((d2, f2), dim=1)))
    cw, co = _head_conv(generator.conv_weight), _head_conv(generator.conv_offset)
    if (generator.kernel_size != 3 or cw.kernel_size != (1, 1) or co.kernel_size != (1, 1) or cw.out_channels != 9
            or co.out_channels != 16 or cw.bias is None or co.bias is None):
        raise NotImplementedError("generator_postprocess needs the reference's 3x3 window and biased 1x1 output "
                                  "convolutions (C -> 9 and C -> 16)")
    conv_w = torch.cat((cw.weight.flatten(1), co.weight.flatten(1)), dim=0)
    conv_b = torch.cat((cw.bias, co.bias))
    residual = postprocessor.residual if hasattr(postprocessor, "residual") else postprocessor.dkn_residual
    mode = NORM_RESIDUAL if residual else NORM_SUM
    init = dem if init_dem is None else init_dem
    return F.gen_propagate(init, feature, conv_w, conv_b, postprocessor.w, postprocessor.b, mode,
                           float(getattr(postprocessor, "scale", 1.0)))


def _head_conv(m: nn.Module) -> nn.Conv2d:
    """The 1x1 convolution inside one of the Generator's two heads: spn.py:41-52 wraps it (Sequential(Conv2d, Sigmoid) /
    Basic2d with .conv = Sequential(Conv2d)), LRRU.py:219-224 uses plain nn.Conv2d modules."""
    if isinstance(m, nn.Conv2d):
        return m
    if hasattr(m, "conv"):
        m = m.conv
    if isinstance(m, nn.Sequential) and len(m) > 0 and isinstance(m[0], nn.Conv2d):
        return m[0]
    if isinstance(m, nn.Conv2d):
        return m
    raise NotImplementedError(f"generator_postprocess: no 1x1 convolution found in {type(m).__name__}")


class Post_process_deconv(nn.Module, ABC):
    """models/LRRU.py:250-298 (`args` needs .kernel_size and .dkn_residual)."""

    def __init__(self, args):
        super().__init__()
        _require_k3(args.kernel_size)
        self.dkn_residual = args.dkn_residual
        self.w = nn.Parameter(torch.ones((1, 1, args.kernel_size, args.kernel_size)))
        self.b = nn.Parameter(torch.zeros(1))
        self.stride = (1, 1)
        self.padding = ((args.kernel_size - 1) // 2, (args.kernel_size - 1) // 2)
        self.dilation = (1, 1)
        self.deformable_groups = 1
        self.im2col_step = 64

    def forward(self, depth, weight, offset):
        mode = NORM_RESIDUAL if self.dkn_residual else NORM_SUM
        return F.propagate(depth, weight, offset, self.w, self.b, mode, 1.0)


class NLSPN(nn.Module):
    """models/components/nlspn.py:8-235.  The 3x3 guidance conv stays a cuDNN conv (it
    is the producer, like spn.Generator); everything after it - offset packing,
    tanh/gamma scaling, confidence gating, abs-sum normalisation, centre weight and the
    prop_time-step loop - runs in the CUDA library."""

    def __init__(self, args, ch_g, ch_f, k_g, k_f):
        super().__init__()
        assert ch_f == 1, "only tested with ch_f == 1 but {}".format(ch_f)
        assert (k_g % 2) == 1, "only odd kernel is supported but k_g = {}".format(k_g)
        pad_g = int((k_g - 1) / 2)
        assert (k_f % 2) == 1, "only odd kernel is supported but k_f = {}".format(k_f)
        pad_f = int((k_f - 1) / 2)
        _require_k3(k_f)

        self.args = args
        self.prop_time = self.args.prop_time
        self.affinity = self.args.affinity
        self.ch_g, self.ch_f, self.k_g, self.k_f = ch_g, ch_f, k_g, k_f
        self.num = self.k_f * self.k_f - 1
        self.idx_ref = self.num // 2

        if self.affinity in ["AS", "ASS", "TC", "TGASS"]:
            self.conv_offset_aff = nn.Conv2d(self.ch_g, 3 * self.num, kernel_size=self.k_g, stride=1,
                                             padding=pad_g, bias=True)
            self.conv_offset_aff.weight.data.zero_()
            self.conv_offset_aff.bias.data.zero_()
            if self.affinity == "TC":
                self.aff_scale_const = nn.Parameter(self.num * torch.ones(1))
                self.aff_scale_const.requires_grad = False
            elif self.affinity == "TGASS":
                self.aff_scale_const = nn.Parameter(self.args.affinity_gamma * self.num * torch.ones(1))
            else:
                self.aff_scale_const = nn.Parameter(torch.ones(1))
                self.aff_scale_const.requires_grad = False
        else:
            raise NotImplementedError

        # frozen gather parameters kept for state_dict compatibility (nlspn.py:61-68)
        self.w = nn.Parameter(torch.ones((self.ch_f, 1, self.k_f, self.k_f)))
        self.b = nn.Parameter(torch.zeros(self.ch_f))
        self.w.requires_grad = False
        self.b.requires_grad = False
        self.w_conf = nn.Parameter(torch.ones((1, 1, 1, 1)))
        self.w_conf.requires_grad = False

        self.stride = 1
        self.padding = pad_f
        self.dilation = 1
        self.groups = self.ch_f
        self.deformable_groups = 1
        self.im2col_step = 64

    def _get_offset_affinity(self, guidance, confidence=None, rgb=None):
        offset_aff = self.conv_offset_aff(guidance)
        conf = confidence if self.args.conf_prop else None
        return F.nlspn_affinity(offset_aff, conf, self.aff_scale_const, self.affinity,
                                bool(getattr(self.args, "legacy", False)) and conf is not None)

    def _propagate_once(self, feat, offset, aff):
        return F.iterate(feat, aff, offset, 1)[0]

    def forward(self, feat_init, guidance, confidence=None, feat_fix=None, rgb=None):
        assert self.ch_g == guidance.shape[1]
        assert self.ch_f == feat_init.shape[1]
        if self.args.conf_prop:
            assert confidence is not None
        offset, aff = self._get_offset_affinity(guidance, confidence if self.args.conf_prop else None, rgb)

        mask_fix = None
        fix = None
        if self.args.preserve_input:
            assert feat_init.shape == feat_fix.shape
            mask_fix = torch.sum(feat_fix > 0.0, dim=1, keepdim=True).detach()
            mask_fix = (mask_fix > 0.0).type_as(feat_fix)
            fix = feat_fix

        if self.prop_time < 1:   # nlspn.py:222-235 with an empty loop: the input comes back, list_feat is empty
            return feat_init, [], offset, aff, self.aff_scale_const.data
        feats = F.iterate(feat_init, aff, offset, self.prop_time, fix, mask_fix)
        list_feat = list(feats.unbind(0))
        feat_result = list_feat[-1]
        return feat_result, list_feat, offset, aff, self.aff_scale_const.data
