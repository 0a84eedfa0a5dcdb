"""Row-strip sharding of a large raster across ranks (one process per GPU).

New relative to the reference, which runs `upscale_dem` on one device
(utils/utils.py:1556-1654).  Each rank owns a band of rows of the DEM, the
affinities and the offsets.  The propagation only couples neighbouring bands
through the rows a tap can reach: `halo = ceil(max |row offset|) + 2` rows of the
DEM on each side.  Those rows are exchanged with the two
neighbours by point-to-point send/recv (NCCL over NVLink on the GPU box, gloo in
the CPU tests); no collective touches the bulk data.  For T > 1 (fixed-affinity
loop) the boundary rows of the *feature* are exchanged after every iteration.

The strip kernel forms coordinates from global row indices, so the concatenated
strips equal the unsharded result bit for bit.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def strip_bounds(H: int, world: int, rank: int, halo: int) -> Tuple[int, int, int, int]:
    """(row0, row1, init_row0, init_row1): rank owns rows [row0,row1); its DEM buffer
    spans [init_row0, init_row1) = the band plus `halo` rows clipped to the image."""
    r0 = (H * rank) // world
    r1 = (H * (rank + 1)) // world
    return r0, r1, max(0, r0 - halo), min(H, r1 + halo)


def exchange_halo(band: torch.Tensor, halo: int, rank: int, world: int, group=None) -> torch.Tensor:
    """band [B,1,Hs,W] -> [B,1,top+Hs+bot,W] with `halo` rows from each existing neighbour.
    Works on any backend with send/recv (nccl for CUDA tensors, gloo for CPU tensors)."""
    if halo <= 0 or world == 1:
        return band
    if band.shape[2] < halo:
        raise RuntimeError(f"halo {halo} exceeds the band height {band.shape[2]}: use fewer ranks")
    up, down = rank - 1, rank + 1
    ops, top, bot = [], None, None
    send_top = band[:, :, :halo].contiguous()
    send_bot = band[:, :, -halo:].contiguous()
    if up >= 0:
        top = torch.empty_like(send_top)
        ops += [dist.P2POp(dist.isend, send_top, up, group), dist.P2POp(dist.irecv, top, up, group)]
    if down < world:
        bot = torch.empty_like(send_bot)
        ops += [dist.P2POp(dist.isend, send_bot, down, group), dist.P2POp(dist.irecv, bot, down, group)]
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    parts = [p for p in (top, band, bot) if p is not None]
    return torch.cat(parts, dim=2)


def global_halo(offset_band: torch.Tensor, group=None, absmax_fn=None) -> int:
    """halo rows needed by every rank: ceil(max over ranks of max |row offset|) + 2."""
    if absmax_fn is None:
        from .functional import offset_absmax as absmax_fn
    m = absmax_fn(offset_band)[:1].clone()
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    return int(math.ceil(float(m.item()))) + 2


class StripPropagator:
    """Propagation of one rank's band.  `H_img` is the height of the whole raster."""

    def __init__(self, H_img: int, rank: Optional[int] = None, world: Optional[int] = None, group=None):
        self.H_img = H_img
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.row0, self.row1, _, _ = strip_bounds(H_img, self.world, self.rank, 0)

    def _buffer(self, band, halo):
        buf = exchange_halo(band, halo, self.rank, self.world, self.group)
        init_row0 = self.row0 - (halo if self.rank > 0 else 0)
        return buf, init_row0

    def forward(self, init_band, weight_band, offset_band, w, b, norm_mode, scale=1.0, halo: Optional[int] = None):
        """One application (JSPSR, T = 1): one halo exchange of the DEM, then the strip kernel."""
        from . import functional as F
        if halo is None:
            halo = global_halo(offset_band, self.group)
        buf, init_row0 = self._buffer(init_band, halo)
        status = torch.zeros(1, dtype=torch.int32, device=init_band.device)
        out = F.spn_forward_strip(buf, weight_band, offset_band, w, b, norm_mode, scale, self.H_img, self.row0,
                                  init_row0, status)
        return out, status

    def iterate(self, feat_band, aff_band, offset_band, T: int, halo: Optional[int] = None):
        """T fixed-affinity applications (NLSPN loop): halo exchange of the feature every iteration."""
        from . import functional as F
        if halo is None:
            halo = global_halo(offset_band, self.group)
        status = torch.zeros(1, dtype=torch.int32, device=feat_band.device)
        feats = []
        cur = feat_band
        for _ in range(T):
            buf, init_row0 = self._buffer(cur, halo)
            cur = F.spn_forward_strip(buf, aff_band, offset_band, None, None, F.NORM_NONE, 0.0, self.H_img,
                                      self.row0, init_row0, status)
            feats.append(cur)
        return feats, status
