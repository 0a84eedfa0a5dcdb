"""Autograd-aware Python entry points over the C ABI (include/jspsr_spn.h).

Everything here only validates shapes, hands raw device pointers + the current
CUDA stream to libjspsr_spn.so and wires the results into autograd.  There is no
PyTorch implementation of the arithmetic in this package: CPU tensors are
rejected, a missing library raises.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from ._lib import BF16, BWD_ACCUMULATE, BWD_GEN_PREACT, F32, MIXED, NORM_NONE, NORM_RESIDUAL, NORM_SUM  # noqa: F401

_workspaces = {}
_launches = 0  # kernels of libjspsr_spn.so enqueued through this module (bench.py reports it)


def launch_count() -> int:
    """Kernels of libjspsr_spn.so enqueued so far, through ctypes (this module) and through the C++ extension."""
    e = _lib.ext()
    return _launches + (int(e.launch_count()) if e is not None else 0)


def _count(n: int = 1) -> None:
    global _launches
    _launches += n


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise RuntimeError(f"jspsr_b200: unsupported dtype {t.dtype} (float32 and bfloat16 only)")


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("jspsr_b200 runs on CUDA tensors only (there is no CPU fallback); got a "
                               f"{t.device} tensor")


def _stream_ptr(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _workspace(t: torch.Tensor) -> torch.Tensor:
    """Zero-initialised reduction scratch, one per (device, stream); the kernels leave it zeroed."""
    key = (t.device.index, _stream_ptr(t))
    ws = _workspaces.get(key)
    if ws is None:
        ws = torch.zeros(_lib.lib().jspsr_spn_workspace_bytes(), dtype=torch.uint8, device=t.device)
        _workspaces[key] = ws
    return ws


def _check_shapes(init, weight, offset):
    # mirrors the checks of torchvision.ops.deform_conv2d for this configuration
    if init.dim() != 4 or init.shape[1] != 1:
        raise RuntimeError(f"init must be [B,1,H,W], got {tuple(init.shape)}")
    B, _, H, W = init.shape
    if tuple(weight.shape) != (B, 9, H, W):
        raise RuntimeError(f"mask/weight must be [B,9,H,W] = {(B, 9, H, W)}, got {tuple(weight.shape)}")
    if offset.dim() != 4 or offset.shape[1] % 18 != 0 or tuple(offset.shape) != (B, 18, H, W):
        raise RuntimeError(f"offset must be [B,2*3*3,H,W] = {(B, 18, H, W)}, got {tuple(offset.shape)}")
    if weight.dtype != offset.dtype or not (init.dtype == weight.dtype or _is_mixed(init, weight)):
        raise RuntimeError("init, weight and offset must share one dtype (or: float32 init with bfloat16 weight/offset)")
    return B, H, W


def _is_mixed(init, weight) -> bool:
    """fp32 DEM with bf16 affinities/offsets: what torch.autocast(bfloat16) hands to PostProcessor.forward."""
    return init.dtype == torch.float32 and weight.dtype == torch.bfloat16


def _io_code(init, weight) -> int:
    return MIXED if _is_mixed(init, weight) else _dtype_code(init)


def _w9(w: Optional[torch.Tensor], like: torch.Tensor) -> Optional[torch.Tensor]:
    if w is None:
        return None
    if w.numel() != 9:
        raise RuntimeError(f"only kernel_size 3 is supported (w has {w.numel()} elements)")
    return w.detach().to(device=like.device, dtype=torch.float32).contiguous()


# ---------------------------------------------------------------------------
# raw calls (no autograd)
# ---------------------------------------------------------------------------
def spn_forward(init, weight, offset, w, b, norm_mode: int, scale: float = 1.0) -> torch.Tensor:
    _require_cuda(init, weight, offset, w, b)
    B, H, W = _check_shapes(init, weight, offset)
    init, weight, offset = init.contiguous(), weight.contiguous(), offset.contiguous()
    w9 = _w9(w, init)
    b1 = b.detach().to(device=init.device, dtype=torch.float32).contiguous()
    out = torch.empty_like(init)
    with torch.cuda.device(init.device):
        rc = _lib.lib().jspsr_spn_forward(_ptr(init), _ptr(weight), _ptr(offset), _ptr(w9), _ptr(b1), _ptr(out),
                                          B, H, W, norm_mode, float(scale), _io_code(init, weight), _stream_ptr(init))
    _lib.check(rc, "jspsr_spn_forward")
    _count()
    return out


def spn_backward(grad_out, init, weight, offset, w, norm_mode: int, scale: float = 1.0, need_grad_init=True,
                 need_grad_w=True, accumulate_into=None, gen_preact=False, reducer=None):
    """Returns (grad_init fp32 | None, grad_weight, grad_offset, grad_w [1,1,3,3] | None, grad_b [1] | None).
    `reducer` (peer.PeerGradReducer): grad_w / grad_b are all-reduced over its ranks inside the kernel.
    `accumulate_into=(grad_weight, grad_offset)` adds into existing buffers (fixed-affinity loops).
    `gen_preact`: generator-tail training - grad_weight is [B,25,H,W] (gradients w.r.t. the pre-activations of the
    Generator's two 1x1 convolutions: sigmoid' applied, centre offset pair dropped) and grad_offset is None."""
    _require_cuda(grad_out, init, weight, offset, w)
    B, H, W = _check_shapes(init, weight, offset)
    grad_out = grad_out.to(init.dtype).contiguous()
    init, weight, offset = init.contiguous(), weight.contiguous(), offset.contiguous()
    w9 = _w9(w, init)
    dev = init.device
    grad_init = torch.empty(B, 1, H, W, dtype=torch.float32, device=dev) if need_grad_init else None
    flags = 0
    if gen_preact:
        if need_grad_init or accumulate_into is not None:
            raise RuntimeError("gen_preact excludes grad_init and accumulation")
        grad_weight, grad_offset = torch.empty(B, 25, H, W, dtype=weight.dtype, device=dev), None
        flags |= BWD_GEN_PREACT
    elif accumulate_into is not None:
        grad_weight, grad_offset = accumulate_into
        flags |= BWD_ACCUMULATE
    else:
        grad_weight, grad_offset = torch.empty_like(weight), torch.empty_like(offset)
    grad_w = torch.empty(1, 1, 3, 3, dtype=torch.float32, device=dev) if need_grad_w else None
    grad_b = torch.empty(1, dtype=torch.float32, device=dev) if need_grad_w else None
    ws = _workspace(init) if need_grad_w else None
    with torch.cuda.device(dev):
        rc = _lib.lib().jspsr_spn_backward_reduce(_ptr(grad_out), _ptr(init), _ptr(weight), _ptr(offset), _ptr(w9),
                                                  _ptr(grad_init), _ptr(grad_weight), _ptr(grad_offset), _ptr(grad_w),
                                                  _ptr(grad_b), _ptr(ws), B, H, W, norm_mode, float(scale),
                                                  _io_code(init, weight), flags,
                                                  reducer.address if (reducer is not None and need_grad_w) else None,
                                                  _stream_ptr(init))
    _lib.check(rc, "jspsr_spn_backward")
    _count()
    return grad_init, grad_weight, grad_offset, grad_w, grad_b


def _check_strip(init_buf, weight, offset, out):
    """Shapes / dtypes of a row-strip call: weight [B,9,Hs,W], offset [B,18,Hs,W], init_buf [B,1,rows >= 1,W]; one dtype,
    or the torch.autocast mix (fp32 DEM buffer and output, bf16 weight / offset).  Returns (B, Hs, W, dtype code)."""
    if weight.dim() != 4 or weight.shape[1] != 9:
        raise RuntimeError(f"weight must be [B,9,Hs,W], got {tuple(weight.shape)}")
    B, _, Hs, W = weight.shape
    if tuple(offset.shape) != (B, 18, Hs, W):
        raise RuntimeError(f"offset must be [B,18,Hs,W] = {(B, 18, Hs, W)}, got {tuple(offset.shape)}")
    if init_buf.dim() != 4 or init_buf.shape[0] != B or init_buf.shape[1] != 1 or init_buf.shape[3] != W or init_buf.shape[2] < 1:
        raise RuntimeError(f"init_buf must be [B,1,rows,W] with B = {B}, W = {W}, got {tuple(init_buf.shape)}")
    if weight.dtype != offset.dtype or not (init_buf.dtype == weight.dtype or _is_mixed(init_buf, weight)):
        raise RuntimeError("init_buf, weight and offset must share one dtype (or: float32 init_buf with bfloat16 weight/offset)")
    if out is not None and (tuple(out.shape) != (B, 1, Hs, W) or not out.is_contiguous() or out.dtype != init_buf.dtype):
        raise RuntimeError("out must be a contiguous [B,1,Hs,W] tensor of init_buf's dtype")
    return B, Hs, W, _io_code(init_buf, weight)


def spn_forward_strip(init_buf, weight, offset, w, b, norm_mode, scale, H_img, row0, init_row0, status=None, out=None,
                      strip_peer=None):
    """Row-strip forward: `init_buf` holds rows [init_row0, init_row0+init_buf.shape[2]) of the image.
    `out` (optional, contiguous [B,1,Hs,W], e.g. the interior of the next halo buffer) receives the result.
    `strip_peer` (peer.StripPeerStruct, B = 1): the halo exchange with the neighbouring ranks runs inside the kernel."""
    _require_cuda(init_buf, weight, offset, out, status)
    B, Hs, W, code = _check_strip(init_buf, weight, offset, out)
    init_buf, weight, offset = init_buf.contiguous(), weight.contiguous(), offset.contiguous()
    w9 = _w9(w, weight)
    b1 = None if b is None else b.detach().to(device=weight.device, dtype=torch.float32).contiguous()
    if out is None:
        out = torch.empty(B, 1, Hs, W, dtype=init_buf.dtype, device=weight.device)
    with torch.cuda.device(weight.device):
        if strip_peer is None:
            rc = _lib.lib().jspsr_spn_forward_strip(_ptr(init_buf), _ptr(weight), _ptr(offset), _ptr(w9), _ptr(b1),
                                                    _ptr(out), B, Hs, W, H_img, row0, init_row0, init_buf.shape[2],
                                                    norm_mode, float(scale), code, _ptr(status), _stream_ptr(weight))
        else:
            import ctypes
            if B != 1:
                raise RuntimeError("the fused halo exchange handles one raster per call (B = 1)")
            rc = _lib.lib().jspsr_spn_forward_strip_peer(_ptr(init_buf), _ptr(weight), _ptr(offset), _ptr(w9), _ptr(b1),
                                                         _ptr(out), Hs, W, H_img, row0, init_row0, init_buf.shape[2],
                                                         norm_mode, float(scale), code, _ptr(status),
                                                         ctypes.addressof(strip_peer), _stream_ptr(weight))
    _lib.check(rc, "jspsr_spn_forward_strip")
    _count()
    return out


def strip_halo_push(band, strip_peer) -> None:
    """First generation of a row-strip sequence: this band's own first / last `halo` rows go to the neighbours' halo rows
    (jspsr_strip_halo_push).  band: contiguous [1,1,Hs,W] view of the buffer that will be read with `strip_peer.stamp`."""
    import ctypes
    _require_cuda(band)
    if band.dim() != 4 or band.shape[0] != 1 or band.shape[1] != 1 or not band.is_contiguous():
        raise RuntimeError(f"band must be a contiguous [1,1,Hs,W] tensor, got {tuple(band.shape)}")
    with torch.cuda.device(band.device):
        rc = _lib.lib().jspsr_strip_halo_push(_ptr(band), band.shape[2], band.shape[3], _dtype_code(band),
                                              ctypes.addressof(strip_peer), _stream_ptr(band))
    _lib.check(rc, "jspsr_strip_halo_push")
    _count()


def gen_spn_forward(init, feature, conv_w, conv_b, w, b, norm_mode: int, scale: float = 1.0, want_weight_offset=False):
    """Generator tail (two 1x1 convolutions, sigmoid, zero centre pair: spn.py:41-52,66-73) fused with the
    propagation (spn.py:99-118).  conv_w [25,C] / conv_b [25]: conv_weight rows first, then conv_offset rows.
    Returns out, or (out, weight [B,9,H,W], offset [B,18,H,W]) when `want_weight_offset`."""
    _require_cuda(init, feature, conv_w, conv_b, w, b)
    if init.dim() != 4 or init.shape[1] != 1:
        raise RuntimeError(f"init must be [B,1,H,W], got {tuple(init.shape)}")
    B, _, H, W = init.shape
    if feature.dim() != 4 or feature.shape[0] != B or tuple(feature.shape[2:]) != (H, W):
        raise RuntimeError(f"feature must be [B,C,H,W] with B,H,W = {(B, H, W)}, got {tuple(feature.shape)}")
    C = feature.shape[1]
    if tuple(conv_w.shape) != (25, C) or tuple(conv_b.shape) != (25,):
        raise RuntimeError(f"conv_w must be [25,{C}] and conv_b [25], got {tuple(conv_w.shape)} / {tuple(conv_b.shape)}")
    if init.dtype != torch.float32 or feature.dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError("gen_spn_forward needs a float32 init and a float32 or bfloat16 (torch.autocast) feature")
    init, feature = init.contiguous(), feature.contiguous()
    conv_w = conv_w.detach().to(torch.float32).contiguous()
    conv_b = conv_b.detach().to(torch.float32).contiguous()
    w9 = _w9(w, init)
    b1 = b.detach().to(device=init.device, dtype=torch.float32).contiguous()
    out = torch.empty_like(init)
    weight = offset = None
    if want_weight_offset:
        weight = torch.empty(B, 9, H, W, dtype=feature.dtype, device=init.device)
        offset = torch.empty(B, 18, H, W, dtype=feature.dtype, device=init.device)
    with torch.cuda.device(init.device):
        rc = _lib.lib().jspsr_gen_spn_forward(_ptr(init), _ptr(feature), _ptr(conv_w), _ptr(conv_b), _ptr(w9), _ptr(b1),
                                              _ptr(out), _ptr(weight), _ptr(offset), B, C, H, W, norm_mode, float(scale),
                                              F32 if feature.dtype == torch.float32 else MIXED, _stream_ptr(init))
    _lib.check(rc, "jspsr_gen_spn_forward")
    _count()
    return (out, weight, offset) if want_weight_offset else out


def gen_tail_grad_feature(gz, conv_w) -> torch.Tensor:
    """grad_feature [B,C,H,W] = sum_j gz[:, j] * conv_w[j, :] (gz [B,25,H,W] from spn_backward(gen_preact=True))."""
    _require_cuda(gz, conv_w)
    if gz.dim() != 4 or gz.shape[1] != 25:
        raise RuntimeError(f"gz must be [B,25,H,W], got {tuple(gz.shape)}")
    B, _, H, W = gz.shape
    C = conv_w.shape[1]
    if conv_w.shape[0] != 25:
        raise RuntimeError(f"conv_w must be [25,C], got {tuple(conv_w.shape)}")
    gz = gz.contiguous()
    conv_w = conv_w.detach().to(torch.float32).contiguous()
    out = torch.empty(B, C, H, W, dtype=gz.dtype, device=gz.device)
    with torch.cuda.device(gz.device):
        rc = _lib.lib().jspsr_gen_tail_grad_feature(_ptr(gz), _ptr(conv_w), _ptr(out), B, C, H, W, _dtype_code(gz),
                                                    _stream_ptr(gz))
    _lib.check(rc, "jspsr_gen_tail_grad_feature")
    _count()
    return out


_gen_workspaces = {}
_GEN_WGRAD_LIBRARY = False   # measurements only (bench.py generator_tail_fused): the cuBLAS batched GEMM the kernel replaced


def gen_tail_grad_params(gz, feature, need_w: bool = True, need_b: bool = True):
    """(grad_conv_w [25,C], grad_conv_b [25]) = (sum_{b,y,x} gz[b,j,y,x] * feature[b,c,y,x], sum_{b,y,x} gz[b,j,y,x]):
    the weight / bias gradients of the Generator tail's two 1x1 convolutions (spn.py:41-52) in one pass over gz and
    feature (gen_tail_wgrad.cu: contraction over the pixel index on tcgen05).  fp32 or bf16 tensors, fp32 results."""
    _require_cuda(gz, feature)
    if gz.dim() != 4 or gz.shape[1] != 25:
        raise RuntimeError(f"gz must be [B,25,H,W], got {tuple(gz.shape)}")
    B, _, H, W = gz.shape
    if feature.dim() != 4 or feature.shape[0] != B or tuple(feature.shape[2:]) != (H, W):
        raise RuntimeError(f"feature must be [B,C,H,W] with B, H, W = {(B, H, W)}, got {tuple(feature.shape)}")
    if gz.dtype != feature.dtype or gz.dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError(f"gen_tail_grad_params takes gz and feature in one dtype, float32 or bfloat16; got {gz.dtype}, "
                           f"{feature.dtype}")
    C = feature.shape[1]
    gz, feature = gz.contiguous(), feature.contiguous()
    key = (gz.device.index, _stream_ptr(gz))
    ws = _gen_workspaces.get(key)
    if ws is None:   # zero on first use; the kernel leaves it zeroed
        ws = torch.zeros(_lib.lib().jspsr_gen_tail_workspace_bytes(), dtype=torch.uint8, device=gz.device)
        _gen_workspaces[key] = ws
    gw = torch.empty(25, C, dtype=torch.float32, device=gz.device) if need_w else None
    gb = torch.empty(25, dtype=torch.float32, device=gz.device) if need_b else None
    with torch.cuda.device(gz.device):
        rc = _lib.lib().jspsr_gen_tail_grad_params(_ptr(gz), _ptr(feature), _ptr(gw), _ptr(gb), _ptr(ws), B, C, H, W,
                                                   _dtype_code(gz), _stream_ptr(gz))
    _lib.check(rc, "jspsr_gen_tail_grad_params")
    _count()
    return gw, gb


def offset_absmax(offset: torch.Tensor) -> torch.Tensor:
    """[max |row offset|, max |col offset|] as a 2-element fp32 device tensor."""
    _require_cuda(offset)
    B, C, H, W = offset.shape
    if C != 18:
        raise RuntimeError("offset must have 18 channels")
    out = torch.zeros(2, dtype=torch.float32, device=offset.device)
    offset = offset.contiguous()
    with torch.cuda.device(offset.device):
        rc = _lib.lib().jspsr_spn_offset_absmax(_ptr(offset), B, H, W, _dtype_code(offset), _ptr(out),
                                                _stream_ptr(offset))
    _lib.check(rc, "jspsr_spn_offset_absmax")
    _count()
    return out


def preserve_blend(feat: torch.Tensor, fix: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The input-preservation blend the LRRU cascade runs between its stages (models/LRRU.py:447-451 and the three
    repeats below it) for a single-channel `fix` (d_clear), in one pass:
        mask = (torch.sum(fix > 0.0, dim=1, keepdim=True) > 0.0).type_as(fix);  (1.0 - mask) * feat + mask * fix
    Same roundings as the torch expression.  No autograd: the reference detaches the result (`output.detach()`).
    `out` may be `feat` itself."""
    _require_cuda(feat, fix, out)
    if feat.shape != fix.shape or feat.dim() != 4 or feat.shape[1] != 1:
        raise RuntimeError(f"preserve_blend takes two [B,1,H,W] tensors, got {tuple(feat.shape)} and {tuple(fix.shape)}")
    if feat.dtype != fix.dtype or (out is not None and (out.dtype != feat.dtype or out.shape != feat.shape)):
        raise RuntimeError("preserve_blend: feat, fix and out must share one dtype (float32 or bfloat16) and shape")
    feat, fix = feat.detach().contiguous(), fix.detach().contiguous()
    if out is None:
        out = torch.empty_like(feat)
    elif not out.is_contiguous():
        raise RuntimeError("preserve_blend: out must be contiguous")
    with torch.cuda.device(feat.device):
        rc = _lib.lib().jspsr_preserve_blend(_ptr(feat), _ptr(fix), _ptr(out), feat.numel(), _dtype_code(feat),
                                             _stream_ptr(feat))
    _lib.check(rc, "jspsr_preserve_blend")
    _count()
    return out


def spn_iterate(feat_init, aff, offset, T: int, feat_fix=None, mask_fix=None) -> torch.Tensor:
    """T fixed-affinity applications; returns all intermediates [T,B,1,H,W].  One dtype for all three tensors
    (jspsr_spn_iterate has no mixed mode; `iterate` promotes autocast's mix first)."""
    _require_cuda(feat_init, aff, offset, feat_fix, mask_fix)
    B, H, W = _check_shapes(feat_init, aff, offset)
    if not (feat_init.dtype == aff.dtype == offset.dtype):
        raise RuntimeError("spn_iterate needs feat_init, aff and offset in one dtype (float32 or bfloat16); "
                           f"got {feat_init.dtype}, {aff.dtype}, {offset.dtype}")
    if T < 1:
        raise RuntimeError(f"spn_iterate: T = {T} must be positive (NLSPN.forward handles prop_time = 0 itself)")
    feat_init, aff, offset = feat_init.contiguous(), aff.contiguous(), offset.contiguous()
    out = torch.empty((T, B, 1, H, W), dtype=feat_init.dtype, device=feat_init.device)
    scratch = None
    if feat_fix is not None:
        feat_fix = feat_fix.to(feat_init.dtype).contiguous()
        mask_fix = mask_fix.to(torch.float32).contiguous()
        scratch = torch.empty_like(feat_init)
    with torch.cuda.device(feat_init.device):
        rc = _lib.lib().jspsr_spn_iterate(_ptr(feat_init), _ptr(aff), _ptr(offset), _ptr(feat_fix), _ptr(mask_fix),
                                          _ptr(out), _ptr(scratch), B, H, W, T, _dtype_code(feat_init),
                                          _stream_ptr(feat_init))
    _lib.check(rc, "jspsr_spn_iterate")
    _count(T * (2 if feat_fix is not None else 1))
    return out


def nlspn_affinity_forward(conv_out, confidence, aff_scale_const, affinity: str, legacy: bool = False):
    _require_cuda(conv_out, confidence, aff_scale_const)
    B, C, H, W = conv_out.shape
    if C != 24:
        raise RuntimeError(f"conv_offset_aff output must have 24 channels (k_f = 3), got {C}")
    conv_out = conv_out.contiguous()
    if confidence is not None:
        if tuple(confidence.shape) != (B, 1, H, W):
            raise RuntimeError(f"confidence must be [B,1,H,W], got {tuple(confidence.shape)}")
        confidence = confidence.to(conv_out.dtype).contiguous()
    gamma = aff_scale_const.detach().to(device=conv_out.device, dtype=torch.float32).contiguous()
    offset = torch.empty(B, 18, H, W, dtype=conv_out.dtype, device=conv_out.device)
    aff = torch.empty(B, 9, H, W, dtype=conv_out.dtype, device=conv_out.device)
    with torch.cuda.device(conv_out.device):
        rc = _lib.lib().jspsr_nlspn_affinity_forward(_ptr(conv_out), _ptr(confidence), _ptr(gamma), _ptr(offset),
                                                     _ptr(aff), B, H, W, _lib.AFFINITY[affinity], int(bool(legacy)),
                                                     _dtype_code(conv_out), _stream_ptr(conv_out))
    _lib.check(rc, "jspsr_nlspn_affinity_forward")
    _count()
    return offset, aff


def nlspn_affinity_backward(grad_offset, grad_aff, conv_out, confidence, aff_scale_const, affinity: str,
                            need_grad_conf=True, need_grad_scale=True):
    _require_cuda(grad_offset, grad_aff, conv_out, confidence)
    B, _, H, W = conv_out.shape
    dev = conv_out.device
    conv_out = conv_out.contiguous()
    grad_offset = grad_offset.to(conv_out.dtype).contiguous()
    grad_aff = grad_aff.to(conv_out.dtype).contiguous()
    if confidence is not None:
        confidence = confidence.to(conv_out.dtype).contiguous()
    gamma = aff_scale_const.detach().to(device=dev, dtype=torch.float32).contiguous()
    grad_conv = torch.empty_like(conv_out)
    need_grad_conf = need_grad_conf and confidence is not None
    grad_conf = torch.empty(B, 1, H, W, dtype=torch.float32, device=dev) if need_grad_conf else None
    need_grad_scale = need_grad_scale and affinity == "TGASS"
    grad_scale = torch.empty(1, dtype=torch.float32, device=dev) if need_grad_scale else None
    ws = _workspace(conv_out) if need_grad_scale else None
    with torch.cuda.device(dev):
        rc = _lib.lib().jspsr_nlspn_affinity_backward(_ptr(grad_offset), _ptr(grad_aff), _ptr(conv_out),
                                                      _ptr(confidence), _ptr(gamma), _ptr(grad_conv), _ptr(grad_conf),
                                                      _ptr(grad_scale), _ptr(ws), B, H, W, _lib.AFFINITY[affinity],
                                                      _dtype_code(conv_out), _stream_ptr(conv_out))
    _lib.check(rc, "jspsr_nlspn_affinity_backward")
    _count()
    return grad_conv, grad_conf, grad_scale


# ---------------------------------------------------------------------------
# autograd Functions
# ---------------------------------------------------------------------------
class _Propagate(torch.autograd.Function):
    """normalise -> deformable 3x3 gather -> (+ scale*init): one kernel each way."""

    @staticmethod
    def forward(ctx, init, weight, offset, w, b, norm_mode, scale, reducer=None):
        ctx.save_for_backward(init, weight, offset, w)
        ctx.norm_mode, ctx.scale, ctx.reducer = norm_mode, scale, reducer
        return spn_forward(init, weight, offset, w, b, norm_mode, scale)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        init, weight, offset, w = ctx.saved_tensors
        need_init = ctx.needs_input_grad[0]
        need_w = ctx.needs_input_grad[3] or ctx.needs_input_grad[4]
        gi, gwt, goff, gw, gb = spn_backward(grad_out, init, weight, offset, w, ctx.norm_mode, ctx.scale,
                                             need_grad_init=need_init, need_grad_w=need_w, reducer=ctx.reducer)
        if gi is not None:
            gi = gi.to(init.dtype)
        if gw is not None:
            gw, gb = gw.to(w.dtype).reshape(w.shape), gb.to(w.dtype)
        return (gi, gwt if ctx.needs_input_grad[1] else None, goff if ctx.needs_input_grad[2] else None,
                gw if ctx.needs_input_grad[3] else None, gb if ctx.needs_input_grad[4] else None, None, None, None)


class _GenPropagate(torch.autograd.Function):
    """Generator tail + propagation.  Forward: one fused kernel that also materialises weight/offset when a
    gradient is needed.  Backward: the fused propagation backward kernel, then the two tensor-core kernels of the
    1x1 convolutions' gradients (feature: gen_tail_backward.cu; weights + biases: gen_tail_wgrad.cu)."""

    @staticmethod
    def forward(ctx, init, feature, conv_w, conv_b, w, b, norm_mode, scale):
        need = any(ctx.needs_input_grad)
        if not need:
            return gen_spn_forward(init, feature, conv_w, conv_b, w, b, norm_mode, scale)
        out, weight, offset = gen_spn_forward(init, feature, conv_w, conv_b, w, b, norm_mode, scale, True)
        ctx.save_for_backward(init, feature, conv_w, weight, offset, w)
        ctx.norm_mode, ctx.scale = norm_mode, scale
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        init, feature, conv_w, weight, offset, w = ctx.saved_tensors
        need_init = ctx.needs_input_grad[0]
        need_w = ctx.needs_input_grad[4] or ctx.needs_input_grad[5]
        B, C, H, W = feature.shape
        if not need_init:
            # the fused backward writes the pre-activation gradients [B,25,H,W] itself (sigmoid', no centre pair)
            gi, gz4, _, gw, gb = spn_backward(grad_out, init, weight, offset, w, ctx.norm_mode, ctx.scale,
                                              need_grad_init=False, need_grad_w=need_w, gen_preact=True)
            gz = gz4.view(B, 25, H * W)
        else:
            gi, gwt, goff, gw, gb = spn_backward(grad_out, init, weight, offset, w, ctx.norm_mode, ctx.scale,
                                                 need_grad_init=True, need_grad_w=need_w)
            gz = torch.empty(B, 25, H * W, dtype=gwt.dtype, device=gwt.device)
            gz4 = gz.view(B, 25, H, W)
            torch.mul(gwt, weight * (1.0 - weight), out=gz4[:, :9])
            gz4[:, 9:17].copy_(goff[:, :8])
            gz4[:, 17:].copy_(goff[:, 10:])
        # the 1x1-convolution parameter gradients: one pass over gz and the feature, contraction over the pixel index on
        # tcgen05 (gen_tail_wgrad.cu)
        g_conv_b = g_conv_w = g_feat = None
        if ctx.needs_input_grad[2] or ctx.needs_input_grad[3]:
            if _GEN_WGRAD_LIBRARY:   # measurements only: the cuBLAS batched GEMM + reduction pass the kernel replaced
                fview = feature.contiguous().view(B, C, H * W)
                g_conv_b = gz.sum(dim=2, dtype=torch.float32).sum(dim=0).to(conv_w.dtype)
                g_conv_w = torch.bmm(gz, fview.transpose(1, 2)).sum(dim=0, dtype=torch.float32).to(conv_w.dtype)
            else:
                f_ = feature if feature.dtype == gz.dtype else feature.to(gz.dtype)
                g_conv_w, g_conv_b = gen_tail_grad_params(gz.view(B, 25, H, W), f_, ctx.needs_input_grad[2],
                                                          ctx.needs_input_grad[3])
                g_conv_w = None if g_conv_w is None else g_conv_w.to(conv_w.dtype)
                g_conv_b = None if g_conv_b is None else g_conv_b.to(conv_w.dtype)
        if ctx.needs_input_grad[1]:   # [B,25,HW] x [25,C] -> [B,C,HW]: tensor-core kernel (gen_tail_backward.cu)
            g_feat = gen_tail_grad_feature(gz.view(B, 25, H, W), conv_w)
        if gw is not None:
            gw, gb = gw.to(w.dtype).reshape(w.shape), gb.to(w.dtype)
        return (gi if need_init else None, g_feat, g_conv_w, g_conv_b, gw if ctx.needs_input_grad[4] else None,
                gb if ctx.needs_input_grad[5] else None, None, None)


def gen_propagate(init, feature, conv_w, conv_b, w, b, norm_mode: int, scale: float = 1.0) -> torch.Tensor:
    return _GenPropagate.apply(init, feature, conv_w, conv_b, w, b, norm_mode, scale)


def _common_dtype(init, weight, offset):
    """A uniform bf16 / fp32 set is used as is, and so is fp32 init with bf16 weight/offset (torch.autocast: the
    kernels read that combination directly - torchvision's operator casts everything to fp32 there, which is the
    same arithmetic).  Any other mix is promoted to float32."""
    if init.dtype == weight.dtype == offset.dtype:
        return init, weight, offset
    if weight.dtype == offset.dtype and _is_mixed(init, weight):
        return init, weight, offset
    return init.float(), weight.float(), offset.float()


def propagate(init, weight, offset, w, b, norm_mode: int, scale: float = 1.0, reducer=None) -> torch.Tensor:
    """`reducer` (peer.PeerGradReducer, batch-sharded training): the gradients of w / b come back all-reduced."""
    init, weight, offset = _common_dtype(init, weight, offset)
    if init.shape[0] == 0:  # empty batch: nothing to launch (the reference returns an empty tensor too)
        _check_shapes(init, weight, offset)
        return init.new_empty(init.shape) + 0 * (weight.sum() + offset.sum() + w.sum() + b.sum())
    e = _lib.ext()
    if e is not None:
        # same checks, same kernels; allocation / stream / workspace / autograd bookkeeping in C++ (torch_binding.cpp):
        # ~3x less host time per call, which is what bounds a step at the reference's batch sizes
        _require_cuda(init, weight, offset, w, b)
        _check_shapes(init, weight, offset)
        if w.numel() != 9:
            raise RuntimeError(f"only kernel_size 3 is supported (w has {w.numel()} elements)")
        return e.propagate(init, weight, offset, w, b, int(norm_mode), float(scale),
                           0 if reducer is None else reducer.address)
    return _Propagate.apply(init, weight, offset, w, b, norm_mode, scale, reducer)


def spn_iterate_backward(grad_list, feat_init, list_out, aff, offset, need_grad_feat=True):
    """Backward of `spn_iterate` without feat_fix in T light carry launches + one gradient kernel
    (jspsr_spn_iterate_backward): returns (grad_feat | None, grad_aff, grad_offset).  fp32, T <= 8."""
    _require_cuda(grad_list, feat_init, list_out, aff, offset)
    B, H, W = _check_shapes(feat_init, aff, offset)
    T = list_out.shape[0]
    if tuple(list_out.shape) != (T, B, 1, H, W) or tuple(grad_list.shape) != (T, B, 1, H, W):
        raise RuntimeError(f"grad_list / list_out must be [T,B,1,H,W], got {tuple(grad_list.shape)} / {tuple(list_out.shape)}")
    for t in (grad_list, feat_init, list_out, aff, offset):
        if t.dtype != torch.float32:
            raise RuntimeError("spn_iterate_backward is fp32 only")
    grad_list, feat_init, list_out = grad_list.contiguous(), feat_init.contiguous(), list_out.contiguous()
    aff, offset = aff.contiguous(), offset.contiguous()
    grad_aff, grad_offset = torch.empty_like(aff), torch.empty_like(offset)
    grad_feat = torch.empty_like(feat_init) if need_grad_feat else None
    carry = torch.empty((T, B, 1, H, W), dtype=torch.float32, device=aff.device)   # T - 1 carries + sum_k |a_k|
    with torch.cuda.device(aff.device):
        rc = _lib.lib().jspsr_spn_iterate_backward(_ptr(grad_list), _ptr(feat_init), _ptr(list_out), _ptr(aff), _ptr(offset),
                                                   _ptr(grad_feat), _ptr(grad_aff), _ptr(grad_offset), _ptr(carry), B, H, W, T,
                                                   F32, _stream_ptr(aff))
    _lib.check(rc, "jspsr_spn_iterate_backward")
    _count((T if need_grad_feat else T - 1) + 1)
    return grad_feat, grad_aff, grad_offset


def _iterate_backward_split_ok(feat_init, aff, offset, feat_fix, T) -> bool:
    """The split backward covers the loop CompletionFormer runs (fp32, no preserve_input, T <= 8).  It pays from three
    steps on (tools/iter_bwd_sweep.py, B200, 2048 / 70 / 2 tiles: T = 1: 2.48 vs 1.63 ms, T = 2: a tie, T = 3: 5.05 vs
    5.34 ms, T = 6: 8.6 vs 11.0 ms), so that is where the default takes it; JSPSR_ITER_BWD=split / steps force one form
    (tests compare the two)."""
    import os
    mode = os.environ.get("JSPSR_ITER_BWD", "auto")
    if mode == "steps":
        return False
    covered = feat_fix is None and T <= 8 and feat_init.dtype == aff.dtype == offset.dtype == torch.float32
    return covered and (T >= 3 or mode == "split")


class _Iterate(torch.autograd.Function):
    """NLSPN loop (nlspn.py:222-235): T applications with fixed aff/offset, w=1, b=0."""

    @staticmethod
    def forward(ctx, feat_init, aff, offset, feat_fix, mask_fix, T):
        out = spn_iterate(feat_init, aff, offset, T, feat_fix, mask_fix)
        ctx.save_for_backward(feat_init, aff, offset, out, feat_fix, mask_fix)
        ctx.T = T
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_list):
        feat_init, aff, offset, out, feat_fix, mask_fix = ctx.saved_tensors
        T = ctx.T
        if _iterate_backward_split_ok(feat_init, aff, offset, feat_fix, T) and grad_list.dtype == torch.float32:
            gf, ga, go = spn_iterate_backward(grad_list, feat_init, out, aff, offset, need_grad_feat=ctx.needs_input_grad[0])
            return (gf, ga if ctx.needs_input_grad[1] else None, go if ctx.needs_input_grad[2] else None, None, None, None)
        grad_aff = grad_offset = None
        carry = None  # gradient flowing into step t's output from step t+1
        for t in range(T - 1, -1, -1):
            g = grad_list[t] if carry is None else grad_list[t] + carry.to(grad_list.dtype)
            src = feat_init if t == 0 else out[t - 1]
            if feat_fix is not None:  # the step consumed the blended input (nlspn.py:229)
                src = ((1.0 - mask_fix) * src + mask_fix * feat_fix).to(src.dtype)
            acc = None if grad_aff is None else (grad_aff, grad_offset)
            gi, grad_aff, grad_offset, _, _ = spn_backward(g, src, aff, offset, None, NORM_NONE, 0.0,
                                                           need_grad_init=True, need_grad_w=False, accumulate_into=acc)
            carry = gi if feat_fix is None else gi * (1.0 - mask_fix)
        return (carry.to(feat_init.dtype) if ctx.needs_input_grad[0] else None,
                grad_aff if ctx.needs_input_grad[1] else None,
                grad_offset if ctx.needs_input_grad[2] else None, None, None, None)


def iterate(feat_init, aff, offset, T: int, feat_fix=None, mask_fix=None) -> torch.Tensor:
    """The loop of NLSPN.forward (nlspn.py:222-235).  Mixed dtypes - e.g. the fp32 feature with bf16 affinities / offsets
    that torch.autocast(bfloat16) produces - are promoted to float32 first, which is what torchvision's operator does with
    them in the reference; a uniform bf16 set runs in bf16 I/O."""
    if not (feat_init.dtype == aff.dtype == offset.dtype):
        feat_init, aff, offset = feat_init.float(), aff.float(), offset.float()
    return _Iterate.apply(feat_init, aff, offset, feat_fix, mask_fix, T)


class _NlspnAffinity(torch.autograd.Function):
    @staticmethod
    def forward(ctx, conv_out, confidence, aff_scale_const, affinity, legacy):
        offset, aff = nlspn_affinity_forward(conv_out, confidence, aff_scale_const, affinity, legacy)
        ctx.save_for_backward(conv_out, confidence, aff_scale_const)
        ctx.affinity, ctx.legacy = affinity, legacy
        return offset, aff

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_offset, grad_aff):
        if ctx.legacy:
            raise RuntimeError("NLSPN legacy mode is inference-only (the reference mutates its offsets in place, "
                               "models/components/nlspn.py:118-128, which autograd rejects there too)")
        conv_out, confidence, gamma = ctx.saved_tensors
        gc, gconf, gscale = nlspn_affinity_backward(grad_offset, grad_aff, conv_out, confidence, gamma, ctx.affinity,
                                                    need_grad_conf=ctx.needs_input_grad[1],
                                                    need_grad_scale=ctx.needs_input_grad[2])
        if gconf is not None:
            gconf = gconf.to(confidence.dtype)
        if gscale is not None:
            gscale = gscale.to(gamma.dtype).reshape(gamma.shape)
        return gc, gconf, gscale, None, None


def nlspn_affinity(conv_out, confidence, aff_scale_const, affinity: str, legacy: bool = False):
    return _NlspnAffinity.apply(conv_out, confidence, aff_scale_const, affinity, legacy)
