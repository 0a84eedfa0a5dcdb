"""Row-strip sharding of a large raster across ranks (one process per GPU).

New relative to the reference, which runs `upscale_dem` on one device
(utils/utils.py:1556-1654).  Each rank owns a band of rows of the DEM, the
affinities and the offsets.  The propagation only couples neighbouring bands
through the rows a tap can reach: `halo = ceil(max |row offset|) + 2` rows of the
DEM on each side.  Those rows are exchanged with the two
neighbours; no collective touches the bulk data.  For T > 1 (fixed-affinity loop) the
boundary rows of the *feature* are exchanged after every iteration.  Two transports:

* `PeerHaloRing` + `StripPropagator.forward_peer / iterate_peer` (GPU box): the exchange
  is FUSED into the propagation kernel (include/jspsr_peer.h) - its edge CTAs store their
  rows straight into the neighbours' next DEM buffer over NVLink and raise a flag there,
  only the edge CTAs of the next application wait for it, the interior of the band is
  computed while the boundary rows travel.  No NCCL, no host synchronisation, no copies.
* point-to-point send/recv (`exchange_halo`, `HaloBuffer.exchange`: NCCL, or gloo in the CPU
  tests) in front of the kernel: the portable form, kept as the cross-check.

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
    m = torch.where(torch.isfinite(m), m, torch.full_like(m, float("inf")))   # NaN does not survive a MAX reduction
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    v = float(m.item())
    if not math.isfinite(v):
        raise RuntimeError("global_halo: the row offsets contain inf or NaN, no finite halo covers them "
                           "(pass an explicit `halo`; taps beyond it raise the kernel's status flag)")
    return int(math.ceil(v)) + 2


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


class PeerHaloRing:
    """`n_buf` DEM buffers of one rank's band ([top + rows + bot, W] each, halo room included) plus the flag block, in
    peer memory, with the two neighbours' rings mapped into this process.  Generations (stamps) count the buffers of a
    sequence: a forward reads generation g and produces g + 1; see include/jspsr_peer.h for the protocol."""

    def __init__(self, rows: int, W: int, halo: int, n_buf: int = 2, dtype=torch.float32, group=None,
                 rank: Optional[int] = None, world: Optional[int] = None, device=None):
        from . import peer
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        if halo < 1 or rows < halo:
            raise RuntimeError(f"halo {halo} must be in [1, band height {rows}]: use fewer ranks")
        if n_buf < 2:
            raise RuntimeError("a ring needs at least two buffers")
        self.rows, self.W, self.halo, self.n_buf, self.dtype = rows, W, halo, n_buf, dtype
        self.top = halo if self.rank > 0 else 0
        self.bot = halo if self.rank < self.world - 1 else 0
        self.es = torch.empty((), dtype=dtype).element_size()
        self.buf_bytes = -(-((self.top + rows + self.bot) * W * self.es) // 256) * 256
        self.flag_bytes = 256
        self._peer = peer
        self.cur = 0          # buffer holding the band that the next application reads
        self.gen = 0          # last generation stamp used
        if self.world == 1:   # a single strip has no neighbours: plain device buffers, no process group needed
            dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
            self.mem = None
            self.status = torch.zeros(1, dtype=torch.int32, device=dev)
            self._bufs = [torch.empty(1, 1, rows, W, dtype=dtype, device=dev) for _ in range(n_buf)]
            return
        geo = [None] * self.world
        dist.all_gather_object(geo, (self.top, rows, self.bot, self.buf_bytes, W, halo, n_buf), group=group)
        if any(g[4:] != (W, halo, n_buf) for g in geo):
            raise RuntimeError("every rank must build its PeerHaloRing with the same W, halo and n_buf")
        self._geo = geo
        nb = [r for r in (self.rank - 1, self.rank + 1) if 0 <= r < self.world]
        self.mem = peer.PeerMemory(self.flag_bytes + n_buf * self.buf_bytes, group, only=nb, device=device)
        self.status = torch.zeros(1, dtype=torch.int32, device=self.mem.device)
        self._bufs = [self.mem.tensor(dtype, self.flag_bytes + k * self.buf_bytes, (self.top + rows + self.bot) * W)
                      .view(1, 1, self.top + rows + self.bot, W) for k in range(n_buf)]

    def buf(self, k: int) -> torch.Tensor:
        return self._bufs[k]

    def interior(self, k: int) -> torch.Tensor:
        return self._bufs[k][:, :, self.top:self.top + self.rows]

    def load(self, band: torch.Tensor) -> None:
        """Copy a [1,1,rows,W] band into the buffer the next application reads (outside any hot loop)."""
        self.interior(self.cur).copy_(band)

    def _halo_dst(self, k):
        """Addresses, in the neighbours' buffer k, of the halo rows this rank fills (None where there is no neighbour)."""
        up = dn = None
        if self.rank > 0:
            top, rows, _, bb = self._geo[self.rank - 1][:4]
            up = self.mem.ptrs[self.rank - 1] + self.flag_bytes + k * bb + (top + rows) * self.W * self.es
        if self.rank < self.world - 1:
            bb = self._geo[self.rank + 1][3]
            dn = self.mem.ptrs[self.rank + 1] + self.flag_bytes + k * bb
        return up, dn

    def strip_peer(self, stamp: int, dst: Optional[int]):
        """jspsr_strip_peer for a call that reads generation `stamp` and pushes its edge rows into the neighbours'
        buffer `dst` (None: signal only)."""
        sp = self._peer.StripPeerStruct()
        if dst is not None:
            sp.up_dst, sp.dn_dst = self._halo_dst(dst)
        sp.up_flags = self.mem.ptrs[self.rank - 1] if self.rank > 0 else None
        sp.dn_flags = self.mem.ptrs[self.rank + 1] if self.rank < self.world - 1 else None
        sp.my_flags = self.mem.local_ptr
        sp.stamp, sp.halo = stamp, self.halo
        return sp

    def close(self) -> None:
        self._bufs = []
        if self.mem is not None:
            self.mem.close()


class StripPropagator:
    """Propagation of one rank's band.  `H_img` is the height of the whole raster."""

    def __init__(self, H_img: int, rank: Optional[int] = None, world: Optional[int] = None, group=None):
        self.H_img = H_img
        self.group = group
        if rank is None or world is None:
            if not dist.is_initialized():
                raise RuntimeError("StripPropagator needs rank= and world= or an initialised process group")
            rank, world = dist.get_rank(group), dist.get_world_size(group)
        self.rank, self.world = rank, world
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

    # ---- exchange fused into the kernel (peer memory over NVLink) ----
    def peer_ring(self, band_rows: int, W: int, halo: int, n_buf: int = 2, dtype=torch.float32) -> PeerHaloRing:
        return PeerHaloRing(band_rows, W, halo, n_buf, dtype, self.group, self.rank, self.world)

    def _push(self, ring: PeerHaloRing) -> int:
        """Start a sequence: this band's edge rows of buffer `ring.cur` go to the neighbours; returns the generation."""
        from . import functional as F
        g = ring.gen + 1
        if self.world > 1:
            F.strip_halo_push(ring.interior(ring.cur), ring.strip_peer(g, ring.cur))
        return g

    def forward_peer(self, ring: PeerHaloRing, weight_band, offset_band, w, b, norm_mode, scale=1.0, out=None):
        """One application (JSPSR, T = 1) of the band loaded in `ring`: returns (out, ring.status)."""
        from . import functional as F
        g = self._push(ring)
        sp = ring.strip_peer(g, None) if self.world > 1 else None
        out = F.spn_forward_strip(ring.buf(ring.cur), weight_band, offset_band, w, b, norm_mode, scale, self.H_img,
                                  self.row0, self.row0 - ring.top, ring.status, out=out, strip_peer=sp)
        ring.gen = g + 1
        return out, ring.status

    def iterate_peer(self, ring: PeerHaloRing, aff_band, offset_band, T: int, keep_all: bool = False):
        """T fixed-affinity applications (NLSPN loop) of the band loaded in `ring`, each one writing the next buffer of the
        ring and its edge rows into the neighbours'.  Returns (list of bands, status): views of the ring's buffers, valid
        until the ring is reused; `keep_all` needs a ring of at least T + 1 buffers."""
        from . import functional as F
        if keep_all and ring.n_buf < T + 1:
            raise RuntimeError(f"keep_all with T = {T} needs a ring of {T + 1} buffers, this one has {ring.n_buf}")
        g = self._push(ring)
        feats = []
        for t in range(T):
            src, dst = ring.cur, (ring.cur + 1) % ring.n_buf
            sp = ring.strip_peer(g, dst) if self.world > 1 else None
            F.spn_forward_strip(ring.buf(src), aff_band, offset_band, None, None, F.NORM_NONE, 0.0, self.H_img, self.row0,
                                self.row0 - ring.top, ring.status, out=ring.interior(dst), strip_peer=sp)
            g += 1
            ring.cur = dst
            if keep_all or t == T - 1:
                feats.append(ring.interior(dst))
        ring.gen = g
        return feats, ring.status

    # ---- exchange by send/recv in front of the kernel (NCCL / gloo) ----
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
