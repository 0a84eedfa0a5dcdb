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


class HaloBuffer:
    """One band of a single raster (B = 1) stored inside a buffer that already has room for the halo rows:
    neighbours' rows are received straight into it and the kernel writes the next iteration's band straight
    into the interior of the other buffer - no concatenation, no copy of the band."""

    def __init__(self, band_rows: int, W: int, halo: int, rank: int, world: int, dtype, device):
        self.top = halo if rank > 0 else 0
        self.bot = halo if rank < world - 1 else 0
        self.halo, self.rank, self.world, self.rows = halo, rank, world, band_rows
        self.buf = torch.empty(1, 1, self.top + band_rows + self.bot, W, dtype=dtype, device=device)

    @property
    def interior(self) -> torch.Tensor:
        return self.buf[:, :, self.top:self.top + self.rows]

    def exchange(self, group=None) -> None:
        """Send this band's first/last `halo` rows to the neighbours, receive theirs into the halo rows."""
        if self.world == 1 or self.halo == 0:
            return
        if self.rows < self.halo:
            raise RuntimeError(f"halo {self.halo} exceeds the band height {self.rows}: use fewer ranks")
        h, t = self.halo, self.top
        ops = []
        if self.rank > 0:
            ops += [dist.P2POp(dist.isend, self.buf[:, :, t:t + h], self.rank - 1, group),
                    dist.P2POp(dist.irecv, self.buf[:, :, :t], self.rank - 1, group)]
        if self.rank < self.world - 1:
            e = t + self.rows
            ops += [dist.P2POp(dist.isend, self.buf[:, :, e - h:e], self.rank + 1, group),
                    dist.P2POp(dist.irecv, self.buf[:, :, e:e + h], self.rank + 1, group)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()


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

    def halo_buffer(self, band: torch.Tensor, halo: int) -> HaloBuffer:
        """Place a [1,1,rows,W] band into a buffer with halo room (one copy, outside any hot loop)."""
        hb = HaloBuffer(band.shape[2], band.shape[3], halo, self.rank, self.world, band.dtype, band.device)
        hb.interior.copy_(band)
        return hb

    def forward(self, init_band, weight_band, offset_band, w, b, norm_mode, scale=1.0, halo: Optional[int] = None):
        """One application (JSPSR, T = 1): one halo exchange of the DEM, then the strip kernel.
        `init_band` is a tensor (exchange + concatenate) or a HaloBuffer already holding the band (zero-copy)."""
        from . import functional as F
        status = torch.zeros(1, dtype=torch.int32, device=weight_band.device)
        if isinstance(init_band, HaloBuffer):
            init_band.exchange(self.group)
            out = F.spn_forward_strip(init_band.buf, weight_band, offset_band, w, b, norm_mode, scale, self.H_img,
                                      self.row0, self.row0 - init_band.top, status)
            return out, status
        if halo is None:
            halo = global_halo(offset_band, self.group)
        buf, init_row0 = self._buffer(init_band, halo)
        out = F.spn_forward_strip(buf, weight_band, offset_band, w, b, norm_mode, scale, self.H_img, self.row0,
                                  init_row0, status)
        return out, status

    def iterate(self, feat_band, aff_band, offset_band, T: int, halo: Optional[int] = None, keep_all: bool = True):
        """T fixed-affinity applications (NLSPN loop): halo exchange of the feature every iteration.
        Single rasters (B = 1) ping-pong between two halo buffers (zero-copy); batches fall back to
        exchange + concatenate.  Returns (list of bands [all T, or just the last], status)."""
        from . import functional as F
        if halo is None:
            halo = global_halo(offset_band, self.group)
        status = torch.zeros(1, dtype=torch.int32, device=feat_band.device)
        feats = []
        if feat_band.shape[0] == 1:
            rows, W = feat_band.shape[2], feat_band.shape[3]
            bufs = [HaloBuffer(rows, W, halo, self.rank, self.world, feat_band.dtype, feat_band.device) for _ in range(2)]
            bufs[0].interior.copy_(feat_band)
            init_row0 = self.row0 - bufs[0].top
            for t in range(T):
                src, dst = bufs[t & 1], bufs[(t + 1) & 1]
                src.exchange(self.group)
                F.spn_forward_strip(src.buf, aff_band, offset_band, None, None, F.NORM_NONE, 0.0, self.H_img,
                                    self.row0, init_row0, status, out=dst.interior)
                if keep_all or t == T - 1:
                    feats.append(dst.interior.clone() if (keep_all and t < T - 1) else dst.interior)
            return feats, status
        cur = feat_band
        for t in range(T):
            buf, init_row0 = self._buffer(cur, halo)
            cur = F.spn_forward_strip(buf, aff_band, offset_band, None, None, F.NORM_NONE, 0.0, self.H_img,
                                      self.row0, init_row0, status)
            if keep_all or t == T - 1:
                feats.append(cur)
        return feats, status
