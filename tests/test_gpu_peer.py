"""The two exchange steps fused into the kernels over peer memory (include/jspsr_peer.h), on 2 GPUs:

* gradient all-reduce inside spn_backward_kernel: sharded gradients == single-process gradients of the concatenated
  batch (DDP semantics around train/train_utils.py:211-219), bit-identical on both ranks, graph-replayable stamps;
* halo exchange inside spn_forward_kernel: row strips == the unsharded call, bit for bit, for T = 1 and the
  fixed-affinity loop (nlspn.py:222-235), TMA and manual staging, fp32 and bf16, and against the send/recv transport.

Skipped on hosts with fewer than 2 GPUs (bench.py runs the same comparisons in-process at every N > 1:
`ddp_parity`, `strips.bitwise_ok`).  Single-GPU parts (one strip through the ring API) run everywhere."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _raster(H, W, seed, device, dtype=torch.float32):
    g = torch.Generator(device="cpu").manual_seed(seed)
    init = torch.rand(1, 1, H, W, generator=g)
    aff = torch.sigmoid(1.5 * torch.randn(1, 9, H, W, generator=g))
    off = (2.0 * torch.randn(1, 18, H, W, generator=g)).clamp_(-6, 6)
    off[:, 8:10] = 0
    return [t.to(device=device, dtype=dtype) for t in (init, aff, off)]


def _strip_checks(rank, world, device):
    from jspsr_b200 import functional as F
    from jspsr_b200.strips import StripPropagator, strip_bounds
    w = torch.full((1, 1, 3, 3), 1.05, device=device)
    b = torch.full((1,), 0.1, device=device)
    results = {}
    for name, (H, W, dtype, env) in {"tma_f32": (300, 512, torch.float32, {}),
                                     "manual_f32": (301, 333, torch.float32, {}),        # W*4 % 16 != 0: no TMA
                                     "narrow_f32": (200, 2048, torch.float32, {}),       # W > 1024: narrow halo variant
                                     "tma_bf16": (300, 512, torch.bfloat16, {})}.items():
        full = _raster(H, W, 5, device, dtype)
        ref1 = F.spn_forward(full[0], full[1], full[2], w, b, 1, 1.0)
        ref = F.spn_iterate(full[0], full[1] * 0.1, full[2], 4)
        r0, r1, _, _ = strip_bounds(H, world, rank, 0)
        band = [t[:, :, r0:r1].contiguous() for t in full]
        sp = StripPropagator(H, rank, world)
        ring = sp.peer_ring(r1 - r0, W, 8, n_buf=5, dtype=dtype)
        ring.load(band[0])
        ok = True
        for _ in range(3):   # repeated sequences reuse the flags / tickets
            out, st = sp.forward_peer(ring, band[1], band[2], w, b, 1, 1.0)
            ok = ok and torch.equal(out, ref1[:, :, r0:r1])
        feats, st = sp.iterate_peer(ring, band[1] * 0.1, band[2], 4, keep_all=True)
        ok = ok and all(torch.equal(f, ref[t][:, :, r0:r1]) for t, f in enumerate(feats))
        ring2 = sp.peer_ring(r1 - r0, W, 8, n_buf=2, dtype=dtype)
        ring2.load(band[0])
        last, st2 = sp.iterate_peer(ring2, band[1] * 0.1, band[2], 4)
        ok = ok and torch.equal(last[-1], ref[3][:, :, r0:r1])
        if world > 1:   # the send/recv transport gives the same bits
            out_sr, st3 = sp.forward(band[0], band[1], band[2], w, b, 1, 1.0, halo=8)
            ok = ok and torch.equal(out_sr, ref1[:, :, r0:r1]) and int(st3.item()) == 0
        ok = ok and int(st.item()) == 0 and int(st2.item()) == 0
        results[name] = bool(ok)
        ring.close()
        ring2.close()
    # a halo that is too small for the offsets raises the status flag instead of returning wrong numbers silently
    if world > 1:
        full = _raster(256, 512, 9, device)
        full[2][:, 0] = 12.0                      # tap 0 reaches 13 rows up
        r0, r1, _, _ = strip_bounds(256, world, rank, 0)
        band = [t[:, :, r0:r1].contiguous() for t in full]
        sp = StripPropagator(256, rank, world)
        ring = sp.peer_ring(r1 - r0, 512, 4, n_buf=2)
        ring.load(band[0])
        _, st = sp.forward_peer(ring, band[1], band[2], w, b, 1, 1.0)
        flag = torch.tensor([int(st.item())], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        results["small_halo_flagged"] = bool(flag.item() == 1)
        ring.close()
    return results


def _reduce_checks(rank, world, device):
    import jspsr_b200
    from jspsr_b200 import functional as F
    from jspsr_b200.peer import PeerGradReducer
    red = PeerGradReducer(average=True)
    per, H, W = 3, 128, 128
    n = per * world
    g = torch.Generator(device="cpu").manual_seed(3)
    dem = torch.rand(n, 1, H, W, generator=g).to(device)
    weight = torch.sigmoid(1.5 * torch.randn(n, 9, H, W, generator=g)).to(device)
    offset = (1.5 * torch.randn(n, 18, H, W, generator=g)).to(device)
    gt = torch.rand(n, 1, H, W, generator=g).to(device)

    def run(sl, reducer, residual=True):
        pp = jspsr_b200.PostProcessor(3, residual, 1.0).to(device).set_grad_reducer(reducer)
        with torch.no_grad():
            pp.w.mul_(1.03)
            pp.b.fill_(0.05)
        wt, of = weight[sl].clone().requires_grad_(), offset[sl].clone().requires_grad_()
        (pp(dem[sl], wt, of) - gt[sl]).square().mean().backward()
        return torch.cat([pp.w.grad.reshape(-1), pp.b.grad.reshape(-1)]), wt.grad, of.grad

    sl = slice(rank * per, (rank + 1) * per)
    out = {}
    for residual in (True, False):
        for rep in range(3):       # odd and even stamps
            flat, gwt, gof = run(sl, red, residual)
        ref, rwt, rof = run(slice(0, n), None, residual)
        rel = lambda a, r: float((a.double() - r.double()).abs().max() / r.double().abs().max())
        both = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(both, flat)
        out[f"residual={residual}"] = {"wb": rel(flat, ref), "weight": rel(gwt / world, rwt[sl]), "offset": rel(gof / world, rof[sl]),
                                       "identical": all(torch.equal(x, flat) for x in both)}
    # sum convention + the raw call inside a CUDA graph (the step stamp is a device-side counter)
    red_sum = PeerGradReducer(average=False)
    w9 = torch.ones(1, 1, 3, 3, device=device)
    gout = torch.ones(per, 1, H, W, device=device)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(2):
            eager = F.spn_backward(gout, dem[sl], weight[sl], offset[sl], w9, 1, 1.0, need_grad_init=False, reducer=red_sum)
    torch.cuda.synchronize()
    dist.barrier()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=s):
        cap = F.spn_backward(gout, dem[sl], weight[sl], offset[sl], w9, 1, 1.0, need_grad_init=False, reducer=red_sum)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    _, _, _, gw_full, gb_full = F.spn_backward(torch.ones(n, 1, H, W, device=device), dem, weight, offset, w9, 1, 1.0,
                                                need_grad_init=False)
    out["graph"] = {"equal_eager": torch.allclose(cap[3], eager[3], rtol=1e-6) and torch.allclose(cap[4], eager[4], rtol=1e-6),
                    "sum_rel": float((cap[3].double() - gw_full.double()).abs().max() / gw_full.double().abs().max()),
                    "b_rel": float((cap[4].double() - gb_full.double()).abs().max() / gb_full.double().abs().max())}
    dist.barrier()
    red.close()
    red_sum.close()
    return out


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    device = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    try:
        res = {"strips": _strip_checks(rank, world, device), "reduce": _reduce_checks(rank, world, device)}
        if rank == 0:
            q.put(res)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_fused_exchange_on_two_gpus():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(res["strips"].values()), res["strips"]
    for name, r in res["reduce"].items():
        if name == "graph":
            assert r["equal_eager"] and r["sum_rel"] < 2e-6 and r["b_rel"] < 2e-6, r
        else:
            assert r["identical"] and r["wb"] < 2e-6 and r["weight"] < 1e-6 and r["offset"] < 1e-6, (name, r)


def test_single_strip_through_the_ring_api():
    """world = 1: the ring API degenerates to plain device buffers and the unsharded kernel (no process group needed)."""
    res = _strip_checks(0, 1, torch.device("cuda", 0))
    assert all(res.values()), res
