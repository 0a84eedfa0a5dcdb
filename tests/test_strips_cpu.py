"""Host-side logic of the row-strip sharding, on CPU with the gloo backend (world_size 2 and 3):
strip bounds, halo exchange between neighbours, halo sizing.  The arithmetic inside a strip is
stood in for by the numpy oracle evaluated on the haloed buffer, which must reproduce the
unsharded oracle result on the rank's own rows."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jspsr_b200.strips import exchange_halo, global_halo, strip_bounds


def test_strip_bounds_cover_the_image():
    for H, world, halo in [(96, 2, 5), (97, 3, 4), (32768, 8, 8), (10, 4, 3)]:
        rows = []
        for r in range(world):
            r0, r1, i0, i1 = strip_bounds(H, world, r, halo)
            assert 0 <= i0 <= r0 < r1 <= i1 <= H
            assert i0 == max(0, r0 - halo) and i1 == min(H, r1 + halo)
            rows += list(range(r0, r1))
        assert rows == list(range(H))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, H, W, halo_in, out_q):
    from oracle import spn_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(3)
        init = rng.random((1, 1, H, W))
        weight = rng.random((1, 9, H, W))
        offset = np.clip(rng.normal(0, 1.2, (1, 18, H, W)), -3.5, 3.5)
        w9, b1 = np.ones(9), 0.1
        full = O.postprocessor_forward(init, weight, offset, w9, b1, O.NORM_RESIDUAL, 1.0)

        r0, r1, _, _ = strip_bounds(H, world, rank, 0)
        band = torch.from_numpy(init[:, :, r0:r1].copy())
        absmax = lambda off: torch.tensor([off[:, 0::2].abs().max(), off[:, 1::2].abs().max()])
        halo = halo_in or global_halo(torch.from_numpy(offset[:, :, r0:r1].copy()), absmax_fn=absmax)
        buf = exchange_halo(band, halo, rank, world)
        _, _, i0, i1 = strip_bounds(H, world, rank, halo)
        assert buf.shape[2] == i1 - i0
        assert np.array_equal(buf.numpy(), init[:, :, i0:i1]), "halo rows are not the neighbours' rows"

        # oracle on the haloed buffer: pad weight/offset rows with zeros outside the band, then cut the band out
        pad_t, pad_b = r0 - i0, i1 - r1
        wpad = np.pad(weight[:, :, r0:r1], ((0, 0), (0, 0), (pad_t, pad_b), (0, 0)))
        opad = np.pad(offset[:, :, r0:r1], ((0, 0), (0, 0), (pad_t, pad_b), (0, 0)))
        part = O.postprocessor_forward(buf.numpy(), wpad, opad, w9, b1, O.NORM_RESIDUAL, 1.0)[:, :, pad_t:pad_t + (r1 - r0)]
        err = float(np.abs(part - full[:, :, r0:r1]).max())
        out_q.put((rank, halo, err))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,halo", [(2, None), (3, None), (2, 6)])
def test_halo_exchange_gloo(world, halo):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 48, 20, halo, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, used_halo, err in results:
        assert used_halo == (halo or 6)   # ceil(3.5) + 2
        assert err < 1e-12, (rank, err)
