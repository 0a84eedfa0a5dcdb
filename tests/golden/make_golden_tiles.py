"""Golden vectors for the tile scheduler + blended merge, produced by the REFERENCE's
own functions.

Run in the build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_tiles.py

utils/utils.py and data/data_utils.py import third-party modules this image does not have (rasterio, rioxarray,
geopandas, affine, natsort, matplotlib ...).  None of them is touched by the functions used here, so they are
replaced by MagicMock stubs and the reference sources are imported UNMODIFIED:

* TileCrop (data/data_utils.py:87-197) cuts seeded samples into tiles (3 x 3, 2 x 2 and 4 x 4 walks);
* add_padding / remove_padding / cal_pad (utils/utils.py:1501-1554) pad seeded images;
* gen_weight_row / gen_weight_col / copyto_add (utils/utils.py:802-913) build the blend weights and perform the
  accumulation of merge_dem (utils/utils.py:916-965).  merge_dem's own body cannot run here (GeoTIFF files through
  rasterio/rioxarray): this script pastes the weighted tiles into their windows in list order through the reference's
  `copyto_add`, which is what `merge_arrays(..., method=copyto_add)` does with them (rasterio.merge.merge).
"""
import importlib
import os
import sys
from unittest.mock import MagicMock

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference(name):
    """Import a reference module unmodified; third-party modules missing from this image become MagicMocks."""
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    stubbed = []
    for _ in range(100):
        try:
            return importlib.import_module(name), stubbed
        except ModuleNotFoundError as e:
            if e.name.split(".")[0] in ("utils", "data", "models", "evaluation", "losses"):
                raise
            sys.modules[e.name] = MagicMock(name=e.name)
            stubbed.append(e.name)
    raise RuntimeError("too many missing modules")


def paste(dest, dest_nodata, new, r0, c0, copyto):
    """rasterio.merge.merge's inner step: the destination window, its nodata mask, the source's mask -> copyto."""
    h, w = new.shape[-2:]
    region = dest[:, r0:r0 + h, c0:c0 + w]
    region_mask = dest_nodata[:, r0:r0 + h, c0:c0 + w]
    copyto(region, new, region_mask, np.zeros_like(region_mask))
    region_mask[...] = False


def reference_merge(U, tiles, border, full):
    """merge_dem (utils/utils.py:916-965) on in-memory tiles [N,k,k]: the reference's weights and copyto_add."""
    n, k, _ = tiles.shape
    n_x = int(round(n ** 0.5))
    b = int(k * border)
    length = k - 2 * b
    out = full - 2 * b
    stride = (out - length) // (n_x - 1)
    weighted = []
    for i in range(n):
        sda = tiles[i:i + 1, b:k - b, b:k - b]                       # clip_box to the buffered footprint
        weighted.append(np.multiply(sda, U.gen_weight_row(sda, i, k)))
    rows = []
    for r in range(n_x):
        dest = np.full((1, length, out), -99999.0, weighted[0].dtype)
        nodata = np.ones(dest.shape, bool)
        for c in range(n_x):
            paste(dest, nodata, weighted[r * n_x + c], 0, stride * c, U.copyto_add)
        rows.append(dest)
    dest = np.full((1, out, out), -99999.0, rows[0].dtype)
    nodata = np.ones(dest.shape, bool)
    for r, xds in enumerate(rows):
        paste(dest, nodata, np.multiply(xds, U.gen_weight_col(xds, r, k)), stride * r, 0, U.copyto_add)
    return dest.squeeze()


def main():
    U, stubs_u = import_reference("utils.utils")
    D, stubs_d = import_reference("data.data_utils")
    print("stubbed third-party modules:", sorted(set(stubs_u + stubs_d)))
    rng = np.random.default_rng(4242)
    res = {}

    # --- TileCrop: 334 -> 9 tiles of 128 is configs r3 (utils/config.py:45-46)
    # (small sizes keep the fixture small; the 334 / 128 geometry itself is recorded in crop334_meta)
    res["crop334_meta"] = np.array([334, 128, *D.TileCrop.get_tile(334, 128)])
    for tag, size, k, n_tile in (("crop70", 70, 32, None), ("crop56", 56, 32, 4), ("crop129", 129, 33, None)):
        img = rng.random((size, size, 3), dtype=np.float32)
        tc = D.TileCrop(crop_size=k, n_tile=n_tile)
        stride, n = D.TileCrop.get_tile(size, k, n_tile)
        tiles = []
        for _ in range(n):
            tiles.append(tc({"image": img})["image"].transpose(2, 0, 1).copy())
        res[f"{tag}_img"] = img
        res[f"{tag}_tiles"] = np.stack(tiles)
        res[f"{tag}_meta"] = np.array([size, k, stride, n])

    # --- mirror border (upscale_dem, utils/utils.py:1557-1580): 50 -> 64 (pad 7), 100 -> 128 (pad 14); 334 -> 512 is pad 89
    for tag, size, c in (("pad50", 50, 1), ("pad100", 100, 3)):
        img = rng.random((size, size, c), dtype=np.float32)
        pad = U.cal_pad(img)
        padded = U.add_padding(img, pad)
        assert np.array_equal(U.remove_padding(padded, pad), img)
        res[f"{tag}_img"] = img
        res[f"{tag}_padded"] = padded
        res[f"{tag}_pad"] = np.array([pad])
    assert U.cal_pad(np.zeros((128, 128, 1))) == 0 and U.cal_pad(np.zeros((334, 334, 1))) == 89

    # --- blend weights and merged rasters
    for tag, full, k, n, border in (("merge9", 334, 128, 9, 0.05), ("merge4", 334, 256, 4, 0.05),
                                    ("merge9_b0", 334, 128, 9, 0.0)):      # full is hard-coded to 334 (utils.py:806)
        # elevations on a 1/4 m grid, stored as int16 (tiles = q / 4): keeps the fixture small
        q = rng.integers(-200, 3200, (n, k, k)).astype(np.int16)
        tiles = (q.astype(np.float32) * np.float32(0.25)).astype(np.float32)
        res[f"{tag}_tiles_q"] = q
        merged = reference_merge(U, tiles, border, full)
        assert merged.dtype == np.float64
        # the full raster for the configs' case, a sub-grid (every 3rd row, every 2nd column) for the others
        res[f"{tag}_merged"] = merged if tag == "merge9" else merged[::3, ::2].copy()
        res[f"{tag}_meta"] = np.array([full, k, n, border])
        b = int(k * border)
        sda = tiles[:1, b:k - b, b:k - b]
        res[f"{tag}_wrow"] = np.stack([U.gen_weight_row(sda, i, k)[0] for i in range(n)])       # [n, L] (one row each)
        res[f"{tag}_wcol"] = np.stack([U.gen_weight_col(np.zeros((1, k - 2 * b, full - 2 * b)), i, k)[:, 0]
                                       for i in range(int(round(n ** 0.5)))])
    path = os.path.join(HERE, "tiles_reference.npz")
    np.savez_compressed(path, **res)
    print("wrote", path, {k_: v.shape for k_, v in res.items()})


if __name__ == "__main__":
    main()
