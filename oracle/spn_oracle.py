"""CPU oracle for the JSPSR spatial-propagation hot path (TEST INFRASTRUCTURE ONLY).

This file is a numpy restatement of the arithmetic the reference performs on
its propagation path.  It is the *checker*: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  Nothing under ``jspsr_b200/`` imports it and
the product path raises when its CUDA library is missing.

What it follows (reference tree = xandercai/JSPSR, paths relative to it):

* ``models/components/spn.py:99-118``  PostProcessor.forward  (normalise ->
  deform_conv2d -> + scale*init)
* ``models/LRRU.py:267-298``           Post_process_deconv.forward (same, no scale); ``:447-498`` the four-stage
  cascade with its input-preservation blend (``lrru_preserve_blend``)
* ``models/components/nlspn.py:77-235`` NLSPN affinity front-end + T-step loop
* ``models/components/spn.py:41-52,66-73`` Generator tail (the two 1x1 convolutions that
  produce weight and offset, sigmoid, zero centre pair) - pinned by
  ``tests/golden/make_golden_generator.py``'s fixtures (gen_*.npz)
* the arithmetic itself lives in a THIRD-PARTY dependency that is not vendored
  in the reference: ``torchvision.ops.deform_conv2d`` (DCNv2, modulated).
  Reference pin: torchvision 0.16 / torch 2.1.0 (``ReadMe.md:9-19``,
  ``Dockerfile:1``).  The restatement below follows that operator's published
  CPU algorithm (``deformable_im2col`` -> ``bilinear_interpolate``; backward:
  ``deformable_col2im`` and ``deformable_col2im_coord`` ->
  ``get_coordinate_weight``).

Parity pin: the reference ships no tests or golden vectors, so
the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: the reference's
own modules were imported from /root/reference in the build container and run
on seeded inputs by ``tests/golden/make_golden.py``; the resulting fixtures are
committed under ``tests/golden/`` and ``tests/test_oracle_golden.py`` checks
this file against every one of them (torchvision 0.26.0+cu128 CPU kernels were
the operator implementation underneath).

All functions are dtype-generic (float32 or float64 numpy arrays, NCHW).
"""
from __future__ import annotations

import numpy as np

KS = 3          # kernel size of the propagation window (reference: always 3)
K = KS * KS     # taps
NORM_NONE, NORM_RESIDUAL, NORM_SUM = 0, 1, 2


# --------------------------------------------------------------------------
# bilinear sampling: torchvision `bilinear_interpolate` / `get_coordinate_weight`
# --------------------------------------------------------------------------
def _tap_positions(offset: np.ndarray, k: int):
    """Sample position of tap k for every pixel.

    torchvision deformable_im2col (stride 1, pad 1, dilation 1):
        y = (out_y*1 - 1) + i*1 + offset_h ; x likewise, i = k//3, j = k%3
    offset channel 2k is the row (h) offset, 2k+1 the column (w) offset
    (SURVEY.md §8 a3).  The integer part is formed first and converted to the
    floating type, then the offset is added - keep that order so fp32 rounding
    matches.
    """
    B, _, H, W = offset.shape
    dt = offset.dtype
    ys = (np.arange(H, dtype=np.int64) - 1 + k // KS).astype(dt)[None, :, None]
    xs = (np.arange(W, dtype=np.int64) - 1 + k % KS).astype(dt)[None, None, :]
    h = ys + offset[:, 2 * k]
    w = xs + offset[:, 2 * k + 1]
    return h, w


def _corners(img: np.ndarray, h: np.ndarray, w: np.ndarray):
    """Four neighbour values (zero where the neighbour is outside the image)
    and the fractional parts.  img [B,H,W]; h,w [B,H,W] sample positions."""
    B, H, W = img.shape
    dt = img.dtype
    nan = np.isnan(h) | np.isnan(w)
    hs = np.where(nan, 0, h)
    ws = np.where(nan, 0, w)
    h0f = np.floor(hs)
    w0f = np.floor(ws)
    lh = (h - h0f).astype(dt)     # NaN propagates through here on purpose
    lw = (w - w0f).astype(dt)
    # anything <= -2 or >= H has no valid corner; clip before the int cast
    h0 = np.clip(h0f, -2, H).astype(np.int64)
    w0 = np.clip(w0f, -2, W).astype(np.int64)
    h1, w1 = h0 + 1, w0 + 1
    bidx = np.arange(B)[:, None, None]

    def fetch(hi, wi):
        ok = (hi >= 0) & (hi <= H - 1) & (wi >= 0) & (wi <= W - 1)
        v = img[bidx, np.clip(hi, 0, H - 1), np.clip(wi, 0, W - 1)]
        return np.where(ok, v, np.zeros((), dt))

    return fetch(h0, w0), fetch(h0, w1), fetch(h1, w0), fetch(h1, w1), lh, lw


def bilinear(img: np.ndarray, h: np.ndarray, w: np.ndarray) -> np.ndarray:
    """torchvision `bilinear_interpolate`: zero when the sample lies outside
    (-1,H)x(-1,W); otherwise the 4-corner blend with per-corner validity."""
    B, H, W = img.shape
    v1, v2, v3, v4, lh, lw = _corners(img, h, w)
    hh, hw = 1 - lh, 1 - lw
    val = hh * hw * v1 + hh * lw * v2 + lh * hw * v3 + lh * lw * v4
    outside = (h <= -1) | (h >= H) | (w <= -1) | (w >= W)
    return np.where(outside, np.zeros((), img.dtype), val)


def coordinate_weight(img: np.ndarray, h: np.ndarray, w: np.ndarray):
    """torchvision `get_coordinate_weight`: d(bilinear)/dh and d/dw using
    per-corner validity only (there is NO whole-sample early exit in the
    backward of the upstream operator)."""
    v1, v2, v3, v4, lh, lw = _corners(img, h, w)
    d_h = lw * (v4 - v2) + (1 - lw) * (v3 - v1)
    d_w = lh * (v4 - v3) + (1 - lh) * (v2 - v1)
    return d_h, d_w


# --------------------------------------------------------------------------
# a2: affinity normalisation (spn.py:100-103 / LRRU.py:268-271)
# --------------------------------------------------------------------------
def normalise(weight: np.ndarray, mode: int) -> np.ndarray:
    if mode == NORM_RESIDUAL:
        return weight - weight.mean(axis=1, keepdims=True, dtype=weight.dtype)
    if mode == NORM_SUM:
        return weight / weight.sum(axis=1, keepdims=True, dtype=weight.dtype)
    return weight


# --------------------------------------------------------------------------
# a3: the operator  out = b + sum_k w_k * m_k * bilinear(init, p_k)
# --------------------------------------------------------------------------
def deform_gather(init: np.ndarray, offset: np.ndarray, mask: np.ndarray,
                  w9: np.ndarray, b1: float) -> np.ndarray:
    """Modulated deformable 3x3 gather of a 1-channel map (deform_conv2d with
    C_in = C_out = 1, stride 1, pad 1, dilation 1).
    init [B,1,H,W], offset [B,18,H,W], mask [B,9,H,W], w9 [9] -> [B,1,H,W]"""
    B, _, H, W = init.shape
    img = init[:, 0]
    w9 = np.asarray(w9, dtype=init.dtype).reshape(K)
    acc = np.zeros((B, H, W), dtype=init.dtype)
    for k in range(K):
        h, w = _tap_positions(offset, k)
        acc += w9[k] * (mask[:, k] * bilinear(img, h, w))
    acc += np.asarray(b1, dtype=init.dtype).reshape(())
    return acc[:, None]


def postprocessor_forward(init, weight, offset, w9, b1, mode=NORM_RESIDUAL,
                          scale=1.0):
    """PostProcessor.forward (spn.py:99-118).  `mode` NORM_RESIDUAL is
    residual=True (subtract the tap mean, add scale*init afterwards);
    NORM_SUM is residual=False (divide by the tap sum, no skip).  NORM_NONE is
    the bare operator (NLSPN's propagate step, nlspn.py:177-187)."""
    m = normalise(weight, mode)
    out = deform_gather(init, offset, m, w9, b1)
    if mode == NORM_RESIDUAL:
        out = out + np.asarray(scale, dtype=init.dtype) * init
    return out


def lrru_preserve_blend(x, d_clear):
    """The input-preservation blend between the stages of the LRRU cascade (models/LRRU.py:447-451, repeated at
    460-464, 474-478, 488-492): valid pixels of `d_clear` (any channel > 0) replace the running estimate.
    Pinned by tests/golden/cascade_lrru.npz (the reference Model's own forward, make_golden_lrru.py)."""
    mask = (np.sum(d_clear > 0.0, axis=1, keepdims=True) > 0.0).astype(d_clear.dtype)
    return (np.asarray(1.0, d_clear.dtype) - mask) * x + mask * d_clear


# --------------------------------------------------------------------------
# a5: explicit backward (restates deformable_col2im / col2im_coord + the
#     Jacobian of the normalisation); independent of any autograd engine
# --------------------------------------------------------------------------
def postprocessor_backward(grad_out, init, weight, offset, w9, mode=NORM_RESIDUAL,
                           scale=1.0, need_grad_init=True):
    """Returns dict(grad_init, grad_weight, grad_offset, grad_w, grad_b)."""
    B, _, H, W = init.shape
    dt = init.dtype
    img = init[:, 0]
    g = grad_out[:, 0]
    w9 = np.asarray(w9, dtype=dt).reshape(K)
    m = normalise(weight, mode)

    grad_m = np.zeros_like(weight)
    grad_off = np.zeros_like(offset)
    grad_w = np.zeros(K, dtype=dt)
    grad_init = np.zeros((B, H, W), dtype=dt) if need_grad_init else None
    bidx = np.broadcast_to(np.arange(B)[:, None, None], (B, H, W))

    for k in range(K):
        h, w = _tap_positions(offset, k)
        val = bilinear(img, h, w)
        d_h, d_w = coordinate_weight(img, h, w)
        gk = g * w9[k]                       # d L / d column_k
        grad_w[k] = (g * (m[:, k] * val)).sum(dtype=dt)
        grad_m[:, k] = gk * val
        grad_off[:, 2 * k] = gk * m[:, k] * d_h
        grad_off[:, 2 * k + 1] = gk * m[:, k] * d_w
        if need_grad_init:
            # deformable_col2im: scatter to the (up to) four valid corners
            _, _, _, _, lh, lw = _corners(img, h, w)
            h0 = np.clip(np.floor(np.where(np.isnan(h), 0, h)), -2, H).astype(np.int64)
            w0 = np.clip(np.floor(np.where(np.isnan(w), 0, w)), -2, W).astype(np.int64)
            c = gk * m[:, k]
            for (hi, wi, cw) in ((h0, w0, (1 - lh) * (1 - lw)), (h0, w0 + 1, (1 - lh) * lw),
                                 (h0 + 1, w0, lh * (1 - lw)), (h0 + 1, w0 + 1, lh * lw)):
                ok = (hi >= 0) & (hi <= H - 1) & (wi >= 0) & (wi <= W - 1)
                np.add.at(grad_init, (bidx[ok], hi[ok], wi[ok]), (c * cw)[ok])

    # Jacobian of a2
    if mode == NORM_RESIDUAL:
        grad_weight = grad_m - grad_m.mean(axis=1, keepdims=True, dtype=dt)
    elif mode == NORM_SUM:
        s = weight.sum(axis=1, keepdims=True, dtype=dt)
        grad_weight = (grad_m - (grad_m * m).sum(axis=1, keepdims=True, dtype=dt)) / s
    else:
        grad_weight = grad_m
    if need_grad_init:
        if mode == NORM_RESIDUAL:
            grad_init = grad_init + np.asarray(scale, dtype=dt) * g
        grad_init = grad_init[:, None]
    return dict(grad_init=grad_init, grad_weight=grad_weight, grad_offset=grad_off,
                grad_w=grad_w.reshape(1, 1, KS, KS), grad_b=np.array([g.sum(dtype=dt)], dtype=dt))


# --------------------------------------------------------------------------
# a7: NLSPN affinity front-end (nlspn.py:77-175), after the 3x3 conv
# --------------------------------------------------------------------------
# ---------------------------------------------------------------------------
# Generator tail (models/components/spn.py:41-52, 66-73): the last two layers of
# Generator.forward, applied to `feature` = the output of Generator.block (spn.py:65)
# ---------------------------------------------------------------------------
def generator_tail(feature, conv_weight_w, conv_weight_b, conv_offset_w, conv_offset_b):
    """weight = sigmoid(conv1x1(feature)) [B,9,H,W] (spn.py:41-44,66);
    offset = conv1x1(feature) [B,16,H,W] viewed as 8 (dy,dx) pairs with a zero pair inserted at index 4
    (spn.py:45-52,67-73) -> [B,18,H,W]."""
    dt = feature.dtype
    B, C, H, W = feature.shape
    cw = conv_weight_w.reshape(K, C).astype(dt)
    co = conv_offset_w.reshape(2 * (K - 1), C).astype(dt)
    zw = np.einsum("nc,bchw->bnhw", cw, feature) + conv_weight_b.astype(dt).reshape(1, -1, 1, 1)
    zo = np.einsum("nc,bchw->bnhw", co, feature) + conv_offset_b.astype(dt).reshape(1, -1, 1, 1)
    weight = (1.0 / (1.0 + np.exp(-zw))).astype(dt)
    offset = np.concatenate((zo[:, :K - 1], np.zeros((B, 2, H, W), dtype=dt), zo[:, K - 1:]), axis=1)
    return weight, offset


def generator_tail_backward(grad_weight, grad_offset, feature, weight, conv_weight_w, conv_offset_w):
    """Autograd of generator_tail: gradients w.r.t. feature and the four convolution parameters."""
    dt = feature.dtype
    B, C, H, W = feature.shape
    cw = conv_weight_w.reshape(K, C).astype(dt)
    co = conv_offset_w.reshape(2 * (K - 1), C).astype(dt)
    gzw = grad_weight * weight * (1.0 - weight)
    gzo = np.concatenate((grad_offset[:, :K - 1], grad_offset[:, K + 1:]), axis=1)  # the centre pair has no source
    return dict(
        grad_feature=np.einsum("bnhw,nc->bchw", gzw, cw) + np.einsum("bnhw,nc->bchw", gzo, co),
        grad_conv_weight_w=np.einsum("bnhw,bchw->nc", gzw, feature).reshape(conv_weight_w.shape),
        grad_conv_weight_b=gzw.sum(axis=(0, 2, 3)),
        grad_conv_offset_w=np.einsum("bnhw,bchw->nc", gzo, feature).reshape(conv_offset_w.shape),
        grad_conv_offset_b=gzo.sum(axis=(0, 2, 3)),
    )


def nlspn_offset_affinity(offset_aff, confidence, aff_scale_const, affinity="TGASS",
                          conf_prop=True, legacy=False):
    """offset_aff [B,24,H,W] = output of conv_offset_aff (nlspn.py:81).
    Returns (offset [B,18,H,W], aff [B,9,H,W])."""
    B, C, H, W = offset_aff.shape
    dt = offset_aff.dtype
    num = K - 1
    assert C == 3 * num
    o1, o2, aff = offset_aff[:, :num], offset_aff[:, num:2 * num], offset_aff[:, 2 * num:]
    # cat(o1,o2).view(B,8,2,H,W): pair n = channels (2n, 2n+1) of the 16-ch cat;
    # zero pair inserted at idx_ref = 4 (nlspn.py:85-90)
    cat = np.concatenate([o1, o2], axis=1)
    offset = np.concatenate([cat[:, :num], np.zeros((B, 2, H, W), dt), cat[:, num:]], axis=1)

    if affinity in ("AS", "ASS"):
        pass
    elif affinity == "TC":
        aff = np.tanh(aff / np.asarray(100, dt)) / np.asarray(aff_scale_const, dt)
    elif affinity == "TGASS":
        aff = np.tanh(aff / np.asarray(100, dt)) / (np.asarray(aff_scale_const, dt) + np.asarray(1e-8, dt))
    else:
        raise NotImplementedError(affinity)

    if conf_prop:
        conf = confidence[:, 0]
        confs = []
        for idx in range(K):
            ww, hh = idx % KS, idx // KS
            if ww == 1 and hh == 1:
                continue
            # 1x1 deformable gather, pad 0: position = pixel + offset (nlspn.py:130-139)
            if legacy:
                # nlspn.py:118-128 writes the shift through a detached VIEW of
                # `offset`, i.e. in place: the shifted offsets are also what the
                # propagation loop and the caller see afterwards.
                offset[:, 2 * idx] += np.asarray(hh - 1, dt)
                offset[:, 2 * idx + 1] += np.asarray(ww - 1, dt)
            oh = offset[:, 2 * idx]
            ow = offset[:, 2 * idx + 1]
            ys = np.arange(H, dtype=np.int64).astype(dt)[None, :, None]
            xs = np.arange(W, dtype=np.int64).astype(dt)[None, None, :]
            confs.append(bilinear(conf, ys + oh, xs + ow))
        aff = aff * np.stack(confs, axis=1)

    aff_abs_sum = np.abs(aff).sum(axis=1, keepdims=True, dtype=dt) + np.asarray(1e-4, dt)
    if affinity in ("ASS", "TGASS"):
        aff_abs_sum = np.where(aff_abs_sum < 1.0, np.ones((), dt), aff_abs_sum)
    if affinity in ("AS", "ASS", "TGASS"):
        aff = aff / aff_abs_sum
    aff_ref = 1.0 - aff.sum(axis=1, keepdims=True, dtype=dt)
    aff = np.concatenate([aff[:, :num // 2], aff_ref.astype(dt), aff[:, num // 2:]], axis=1)
    return offset, aff


def nlspn_propagate(feat_init, offset, aff, prop_time, feat_fix=None, preserve_input=False):
    """nlspn.py:216-235: T applications of the bare operator (w=1, b=0,
    mask=aff), every intermediate kept."""
    ones = np.ones(K, dtype=feat_init.dtype)
    feat = feat_init
    feats = []
    if preserve_input:
        mask_fix = ((feat_fix > 0).sum(axis=1, keepdims=True) > 0).astype(feat_init.dtype)
    for _ in range(prop_time):
        if preserve_input:
            feat = (1 - mask_fix) * feat + mask_fix * feat_fix
        feat = deform_gather(feat, offset, aff, ones, 0.0)
        feats.append(feat)
    return feat, feats


def nlspn_propagate_backward(grad_list, feat_init, feats, offset, aff):
    """Gradient of the loop above (no preserve_input) w.r.t. feat_init, aff and offset, given the gradient of every
    step's output: what autograd makes of nlspn.py:222-235 - step t receives grad_list[t] plus what step t + 1 sent
    back through its own input, and the gradients of the shared (aff, offset) add up over the steps.
    Checked against central differences in tests/test_oracle_golden.py."""
    ones = np.ones(K, dtype=feat_init.dtype)
    T = len(feats)
    carry = None
    grad_aff = np.zeros_like(aff)
    grad_offset = np.zeros_like(offset)
    for t in range(T - 1, -1, -1):
        g = grad_list[t] if carry is None else grad_list[t] + carry
        src = feat_init if t == 0 else feats[t - 1]
        r = postprocessor_backward(g, src, aff, offset, ones, NORM_NONE, 0.0, need_grad_init=True)
        carry = r["grad_init"]
        grad_aff = grad_aff + r["grad_weight"]
        grad_offset = grad_offset + r["grad_offset"]
    return carry, grad_aff, grad_offset


# --------------------------------------------------------------------------
# consumer arithmetic used for the end-to-end gate (evaluation/metrics.py:147-199,
# 361-396; data/data_utils.py:441-457)
# --------------------------------------------------------------------------
def descale(data, elev_min, elev_max, elev_log=False):
    if elev_log:
        return np.exp(data * np.log(elev_max - elev_min)) + elev_min
    return data * (elev_max - elev_min) + elev_min


def rmse_mae(pred, gt, border=0.05, value_min=0.0, value_max=1.0, elev_log=False):
    """Per-sample RMSE averaged over samples (MeterRMSE.update is called with
    batch size 1) and the matching MAE."""
    h, w = pred.shape[-2:]
    bh, bw = int(h * border), int(w * border)
    p = np.clip(pred[..., bh:h - bh, bw:w - bw], 0.0, 1.0)
    t = gt[..., bh:h - bh, bw:w - bw]
    p = descale(p, value_min, value_max, elev_log)
    t = descale(t, value_min, value_max, elev_log)
    d = (p - t).reshape(p.shape[0], -1)
    rmse = np.sqrt((d ** 2).sum(axis=1) / d.shape[1])
    mae = np.abs(d).mean(axis=1)
    return float(rmse.mean()), float(mae.mean())
