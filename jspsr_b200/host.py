"""Host-buffer entry point: the call a CPU-side caller makes (tensors in host
memory, result wanted in host memory).  Wraps jspsr_spn_forward_host, which cuts the
batch into chunks and overlaps H2D copy / kernel / D2H copy on internal streams.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

_scratch = {}


def _host_ptr(a):
    if isinstance(a, torch.Tensor):
        if a.is_cuda:
            raise RuntimeError("forward_host takes host tensors; use functional.spn_forward for device tensors")
        return a.data_ptr()
    return a.ctypes.data


def _shape_dtype(a):
    if isinstance(a, torch.Tensor):
        return tuple(a.shape), {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}[a.dtype]
    return a.shape, {np.dtype(np.float32): _lib.F32}[a.dtype]


def default_chunk(B: int, H: int, W: int) -> int:
    # ~32 Mpix-channels per chunk: large enough to hide launch latency, small enough to pipeline
    return max(1, min(B, (1 << 20) // max(1, H * W) or 1))


def forward_host(init, weight, offset, w9, b1, norm_mode: int, scale: float = 1.0, chunk_B: int | None = None,
                 out=None, device: int = 0):
    """init [B,1,H,W], weight [B,9,H,W], offset [B,18,H,W]: C-contiguous numpy arrays or CPU
    torch tensors (pinned memory reaches full PCIe rate).  Returns `out` (same kind as `init`)."""
    (B, one, H, W), dtype = _shape_dtype(init)
    if one != 1 or tuple(weight.shape) != (B, 9, H, W) or tuple(offset.shape) != (B, 18, H, W):
        raise RuntimeError("expected init [B,1,H,W], weight [B,9,H,W], offset [B,18,H,W]")
    for a in (init, weight, offset):
        contiguous = a.is_contiguous() if isinstance(a, torch.Tensor) else a.flags["C_CONTIGUOUS"]
        if not contiguous:
            raise RuntimeError("forward_host needs C-contiguous host buffers")
    if chunk_B is None:
        chunk_B = default_chunk(B, H, W)
    if out is None:
        out = torch.empty_like(init) if isinstance(init, torch.Tensor) else np.empty_like(init)
    w9 = np.ascontiguousarray(np.asarray(w9, dtype=np.float32).reshape(9))
    b1 = np.ascontiguousarray(np.asarray(b1, dtype=np.float32).reshape(1))
    lib = _lib.lib()
    need = lib.jspsr_spn_host_scratch_bytes(chunk_B, H, W, dtype)
    key = (device, need)
    scratch = _scratch.get(key)
    if scratch is None:
        _scratch.clear()
        scratch = torch.empty(need, dtype=torch.uint8, device=f"cuda:{device}")
        _scratch[key] = scratch
    with torch.cuda.device(device):
        rc = lib.jspsr_spn_forward_host(_host_ptr(init), _host_ptr(weight), _host_ptr(offset), w9.ctypes.data,
                                        b1.ctypes.data, _host_ptr(out), B, H, W, norm_mode, float(scale), dtype,
                                        scratch.data_ptr(), need, chunk_B)
    _lib.check(rc, "jspsr_spn_forward_host")
    return out
