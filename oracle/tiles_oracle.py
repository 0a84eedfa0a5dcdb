"""CPU oracle (numpy) for the tile scheduler and the blended merge.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
leg as the checker; the product (jspsr_b200/) never imports it.

What it restates, with the reference lines it follows:

* get_tile / tile_origins  - data/data_utils.py:170-194 (TileCrop.get_tile) and :129-163 (row-major walk of
  the n_x * n_x tiles, stride * row / stride * col origins);
* cal_pad / add_padding / remove_padding - utils/utils.py:1501-1554, INCLUDING the bottom border's one-row
  shift (`img_with_border[-2n-1:-n-1]`, :1517): bottom pad row i mirrors image row h-2-i, not h-1-i;
* ramp / weight_1d - utils/utils.py:802-900 (gen_weight_row / gen_weight_col: `linspace(1, 0, p+2)[1:-1]` over
  the p overlapped pixels; first tile ramps down at its end, last tile ramps up at its start, middle tiles both);
* merge_tiles - utils/utils.py:903-965 (merge_dem with method=copyto_add, the only call site :1272): clip
  `int(width * border)` pixels from every side of every tile (:931-934), multiply by the row weight (float64),
  merge each row of tiles left to right, multiply the merged rows by the column weight, merge top to bottom.
  `rioxarray.merge.merge_arrays` -> `rasterio.merge.merge` is a third-party dependency absent from
  /root/reference and from this image: its published algorithm (paste every source into its window of the
  destination in list order through `copyto(region, new, region_nodata_mask, new_nodata_mask)`) is restated here
  with the reference's own `copyto_add` (:903-913): both valid -> add, only the new one valid -> copy.

Pinned by tests/golden/make_golden_tiles.py against the reference's own functions (imported unmodified, with the
absent third-party modules stubbed): gen_weight_row, gen_weight_col, copyto_add, add_padding, remove_padding,
cal_pad, TileCrop.  The reference only accepts 4 or 9 tiles (n_x = 2, 3); n_x >= 2 follows the same rule
(first / middle / last) and n_x = 1 is a plain crop.
"""
from __future__ import annotations

from math import ceil

import numpy as np


def get_tile(w: int, k: int, n=None):
    """(stride, number of tiles) of a square w x w image cut into k x k tiles (data_utils.py:170-194)."""
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


def tile_origins(w: int, k: int, n=None):
    """Row-major (row0, col0) of every tile, the order TileCrop hands them out (data_utils.py:144-163)."""
    stride, n_tile = get_tile(w, k, n)
    n_x = int(round(n_tile ** 0.5))
    return [(stride * r, stride * c) for r in range(n_x) for c in range(n_x)], stride, n_x


def cal_pad(h: int, w: int) -> int:
    """Mirror-border width that brings an h x w image to the next power of two (utils.py:1536-1554)."""
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


def pad_source_index(i: int, n: int, size: int, bottom_quirk: bool) -> int:
    """Source row/column of padded index i (add_padding, utils.py:1501-1522)."""
    if i < n:
        return n - 1 - i
    if i < n + size:
        return i - n
    j = i - n - size
    return size - 2 - j if bottom_quirk else size - 1 - j


def add_padding(img: np.ndarray, n: int) -> np.ndarray:
    """img [H,W,C] -> [H+2n, W+2n, C] float32 with the reference's mirrored border."""
    h, w, c = img.shape
    rows = [pad_source_index(i, n, h, True) for i in range(h + 2 * n)]
    cols = [pad_source_index(j, n, w, False) for j in range(w + 2 * n)]
    return np.ascontiguousarray(img[np.ix_(rows, cols)]).astype(np.float32)


def remove_padding(img: np.ndarray, pad: int) -> np.ndarray:
    h, w, _ = img.shape
    return img[pad:h - pad, pad:w - pad, :]


def crop_tiles(img: np.ndarray, k: int, n=None, pad: int = 0) -> np.ndarray:
    """img [H,W,C] (square) -> [N, C, k, k]: optional mirror border, then the TileCrop walk, channels first
    (ToTensor's HWC -> CHW, data_utils.py:232)."""
    if pad > 0:
        img = add_padding(img, pad)
    h, w, _ = img.shape
    assert h == w
    origins, _, _ = tile_origins(w, k, n)
    return np.stack([img[r:r + k, c:c + k, :].transpose(2, 0, 1) for r, c in origins])


def ramp(p: int) -> np.ndarray:
    return np.linspace(1, 0, p + 2)[1:-1]


def weight_1d(i: int, n_x: int, length: int, p: int) -> np.ndarray:
    """1-D blend weight of tile i of n_x along one axis (gen_weight_row/col, utils.py:818-845)."""
    w = np.ones(length)
    if n_x == 1 or p <= 0:
        return w
    r = ramp(p)
    if i == 0:
        w[-p:] = r
    elif i == n_x - 1:
        w[-p:] = r
        w = np.flip(w).copy()
    else:
        w[:p] = np.flip(r)
        w[-p:] = r
    return w


def merge_geometry(k: int, border: float, full: int):
    b = int(k * border)                       # utils.py:931-934
    length = k - 2 * b                        # border-cropped tile (116 at k = 128, border = 0.05)
    out = full - (k - length)                 # border-cropped full prediction (322)
    stride, n_tile = get_tile(out, length)    # utils.py:812
    n_x = int(round(n_tile ** 0.5))
    return b, length, out, stride, n_x, length - stride


def merge_tiles(tiles: np.ndarray, border: float = 0.05, full: int = 334) -> np.ndarray:
    """tiles [N, k, k] float32 (row-major n_x * n_x predictions of one `full`-sized sample) -> float64 [out, out]."""
    n, k, _ = tiles.shape
    b, length, out, stride, n_x, p = merge_geometry(k, border, full)
    assert n == n_x * n_x, (n, n_x)
    rows = []
    for r in range(n_x):
        strip = np.zeros((length, out), np.float64)
        filled = np.zeros((length, out), bool)
        for c in range(n_x):
            t = tiles[r * n_x + c, b:b + length, b:b + length]
            new = np.multiply(t, weight_1d(c, n_x, length, p)[None, :])      # float32 * float64 -> float64
            region = strip[:, stride * c:stride * c + length]
            seen = filled[:, stride * c:stride * c + length]
            region[seen] = (region + new)[seen]                               # copyto_add: both valid -> add
            region[~seen] = new[~seen]                                        #             only new valid -> copy
            seen[:] = True
        rows.append(strip)
    merged = np.zeros((out, out), np.float64)
    filled = np.zeros((out, out), bool)
    for r in range(n_x):
        new = rows[r] * weight_1d(r, n_x, length, p)[:, None]
        region = merged[stride * r:stride * r + length]
        seen = filled[stride * r:stride * r + length]
        region[seen] = (region + new)[seen]
        region[~seen] = new[~seen]
        seen[:] = True
    return merged
