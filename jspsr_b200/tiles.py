"""Tile scheduler and blended merge on the GPU over the C ABI in include/jspsr_tiles.h.

Mirrors, on device tensors, what the reference does on numpy arrays / GeoTIFF files on the CPU:

* `get_tile`, `TileCrop`  - data/data_utils.py:87-197 (stride and tile count of the overlapping walk);
* `cal_pad`, `crop_tiles(..., pad=)`, `remove_padding` - utils/utils.py:1501-1554 (upscale_dem's mirrored border);
* `merge_tiles` - utils/utils.py:802-965 (merge_dem with linear-ramp weights and copyto_add).

Host logic here is integer bookkeeping only; pixels are moved by `tiles_crop_kernel` / `tiles_merge_kernel`.
There is no CPU fallback.
"""
from __future__ import annotations

from math import ceil

import torch

from . import _lib
from .functional import _count, _ptr, _require_cuda, _stream_ptr


def get_tile(w: int, k: int, n=None):
    """(stride, number of tiles) - TileCrop.get_tile, data/data_utils.py:170-194 (same assertions)."""
    if n is None:
        n_x = (w - w % k) / k + 1
    else:
        n_x = ceil(n ** 0.5)
    assert n_x % 1 == 0, "cannot divide the image into n_tile tiles, check the input."
    if n_x == 1:
        return 0, 1
    stride = (w - k) / (n_x - 1)
    assert stride % 1 == 0, "no padding for cropping to tile evenly, check the input."
    return int(stride), int(n_x ** 2)


def cal_pad(h: int, w: int) -> int:
    """Border that brings (h, w) to the next power of two - cal_pad, utils/utils.py:1536-1554."""
    if int.bit_count(h) == 1 and int.bit_count(w) == 1:
        return 0
    h_pad = w_pad = 0
    for i in range(1, 10):
        if 2 ** i > h:
            h_pad = (2 ** i - h) // 2
            w_pad = (2 ** i - w) // 2
            break
    assert h_pad == w_pad
    return h_pad


def _f32_cuda(t: torch.Tensor, what: str) -> torch.Tensor:
    _require_cuda(t)
    if t.dtype != torch.float32:
        raise RuntimeError(f"jspsr_b200.tiles: {what} must be float32, got {t.dtype}")
    return t.contiguous()


def crop_tiles(raster: torch.Tensor, k: int, n_tile=None, pad: int = 0, stride=None, grid=None) -> torch.Tensor:
    """raster [C,H,W] -> [N,C,k,k] tiles in TileCrop's row-major order; `pad` > 0 first adds upscale_dem's
    mirrored border.  By default the walk is TileCrop's (square rasters); `stride` + `grid=(n_y, n_x)` give an
    explicit walk for rectangular rasters / strips."""
    raster = _f32_cuda(raster, "raster")
    if raster.dim() != 3:
        raise RuntimeError(f"jspsr_b200.tiles: raster must be [C,H,W], got {tuple(raster.shape)}")
    C, H, W = raster.shape
    if grid is None:
        assert H == W, "TileCrop's walk supports square rasters only; pass stride= and grid= otherwise"
        stride, n = get_tile(W + 2 * pad, k, n_tile)
        n_y = n_x = int(round(n ** 0.5))
    else:
        n_y, n_x = grid
        assert stride is not None
    out = torch.empty(n_y * n_x, C, k, k, dtype=torch.float32, device=raster.device)
    with torch.cuda.device(raster.device):
        _lib.check(_lib.lib().jspsr_tiles_crop(_ptr(raster), _ptr(out), C, H, W, pad, k, stride, n_y, n_x,
                                               _stream_ptr(raster)), "jspsr_tiles_crop")
    _count()
    return out


def add_padding(raster: torch.Tensor, pad: int) -> torch.Tensor:
    """raster [C,H,W] -> [C,H+2*pad,W+2*pad] with upscale_dem's mirrored border (utils/utils.py:1501-1522)."""
    C, H, W = raster.shape
    assert H == W, "add_padding pads square rasters (cal_pad asserts h_pad == w_pad, utils.py:1552)"
    return crop_tiles(raster, H + 2 * pad, pad=pad, stride=0, grid=(1, 1))[0]


def remove_padding(t: torch.Tensor, pad: int) -> torch.Tensor:
    """[..., H, W] -> the window without the border (utils/utils.py:1525-1533); a view, no kernel."""
    return t if pad == 0 else t[..., pad:t.shape[-2] - pad, pad:t.shape[-1] - pad]


def merge_geometry(k: int, border: float, full: int):
    """(crop, L, out, stride, n_x): merge_dem's clip (utils.py:931-934) and gen_weight_row's geometry (:806-814)."""
    crop = int(k * border)
    length = k - 2 * crop
    out = full - (k - length)
    stride, n = get_tile(out, length)
    return crop, length, out, stride, int(round(n ** 0.5))


def merge_tiles(tiles: torch.Tensor, border: float = 0.05, full: int = 334, dtype=torch.float64, stride=None,
                grid=None) -> torch.Tensor:
    """tiles [S,N,k,k] (or [N,k,k] / [N,1,k,k] for one sample) -> [S,out,out] (or [out,out]) blended rasters.
    float64 output is bit-identical to merge_dem's numpy result; float32 rounds it on store."""
    tiles = _f32_cuda(tiles, "tiles")
    # [N,k,k] or the model's output batch [N,1,k,k] = the tiles of ONE sample; [S,N,k,k] = S samples
    single = tiles.dim() == 3 or (tiles.dim() == 4 and tiles.shape[1] == 1)
    if single:
        tiles = tiles.reshape(1, -1, tiles.shape[-2], tiles.shape[-1])
    if tiles.dim() != 4:
        raise RuntimeError(f"jspsr_b200.tiles: tiles must be [N,k,k], [N,1,k,k] or [S,N,k,k], got {tuple(tiles.shape)}")
    S, N, k, k2 = tiles.shape
    assert k == k2, "square tiles only"
    crop = int(k * border)
    if grid is None:
        crop, _, _, stride, n_x = merge_geometry(k, border, full)
        n_y = n_x
    else:
        n_y, n_x = grid
        assert stride is not None
    if n_y * n_x != N:
        raise RuntimeError(f"jspsr_b200.tiles: {N} tiles given, the geometry needs {n_y} x {n_x}")
    if dtype not in (torch.float64, torch.float32):
        raise RuntimeError("jspsr_b200.tiles: merged rasters are float64 (the reference's) or float32")
    L = k - 2 * crop
    out = torch.empty(S, stride * (n_y - 1) + L, stride * (n_x - 1) + L, dtype=dtype, device=tiles.device)
    with torch.cuda.device(tiles.device):
        _lib.check(_lib.lib().jspsr_tiles_merge(_ptr(tiles), _ptr(out), S, n_y, n_x, k, crop, stride,
                                                int(dtype == torch.float64), _stream_ptr(tiles)), "jspsr_tiles_merge")
    # 2 x 2 gather: one launch for the rows owned by one tile row, one for the overlapped rows; else the generic kernel
    _count(2 if (n_y > 1 and 0 < L - stride and L <= 2 * stride) else 1)
    return out[0] if single else out


def tiled_apply(fn, raster: torch.Tensor, k: int, pad: int = 0, border: float = 0.05, n_tile=None, batch=None,
                dtype=torch.float64) -> torch.Tensor:
    """Whole-raster inference as the reference composes it, every stage on the GPU: mirror-pad (`upscale_dem`,
    utils/utils.py:1563-1578) -> overlapping tiles (`TileCrop`, data/data_utils.py:129-163) -> `fn` on batches of
    tiles -> border crop + blended merge (`merge_dem`, utils/utils.py:916-965) -> remove the padding (:1641-1642).

    raster [C,H,H]; `fn` maps tiles [n,C,k,k] to predictions [n,1,k,k] (or [n,k,k]); `batch` tiles per call (default:
    all).  The merge walks the SAME stride / grid the crop used.  Returns [h,h] with h = H - 2*max(0, crop - pad),
    crop = int(k*border): the part of the image the merged raster covers (all of it when pad >= crop)."""
    raster = _f32_cuda(raster, "raster")
    C, H, W = raster.shape
    assert H == W, "TileCrop's walk supports square rasters only"
    stride, n = get_tile(W + 2 * pad, k, n_tile)
    n_x = int(round(n ** 0.5))
    tiles = crop_tiles(raster, k, n_tile, pad=pad)
    outs = []
    step = n if not batch else int(batch)
    for i in range(0, n, step):
        o = fn(tiles[i:i + step])
        if o.dim() == 4:
            if o.shape[1] != 1:
                raise RuntimeError(f"jspsr_b200.tiles: fn must return one channel, got {tuple(o.shape)}")
            o = o[:, 0]
        if tuple(o.shape[-2:]) != (k, k):
            raise RuntimeError(f"jspsr_b200.tiles: fn must keep the tile size {k}, got {tuple(o.shape)}")
        outs.append(o)
    pred = torch.cat(outs) if len(outs) > 1 else outs[0]
    merged = merge_tiles(pred.float(), border, dtype=dtype, stride=stride, grid=(n_x, n_x))
    crop = int(k * border)
    lo = max(0, pad - crop)                       # merged index of image row max(0, crop - pad)
    h = H - 2 * max(0, crop - pad)
    return merged[lo:lo + h, lo:lo + h]
