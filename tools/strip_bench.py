"""Row-strip sharded propagation of one large raster across ranks (torchrun, one process per GPU).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/strip_bench.py \
        [--H 32768 --W 32768 --T 1]

Each rank owns H/N rows of the DEM, affinities and offsets (generated on the device, seeded by global row so the
raster is the same for every N).  T = 1 is JSPSR (one halo exchange of the DEM); T > 1 is the fixed-affinity loop with a
halo exchange of the feature after every iteration (NCCL send/recv between neighbours, nothing else crosses GPUs).
Checks: a small raster against the unsharded single-GPU result (bit for bit) before timing.
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from jspsr_b200 import functional as F
from jspsr_b200.strips import StripPropagator, strip_bounds


def make_rows(r0, r1, W, seed):
    """Deterministic per-row content so any sharding sees the same raster."""
    rows = r1 - r0
    g = torch.Generator(device="cuda").manual_seed(seed)
    # generate in row blocks keyed by global block index (block = 256 rows)
    outs = [[], [], []]
    blk = 256
    b0 = r0 // blk
    b1 = (r1 + blk - 1) // blk
    for b in range(b0, b1):
        g.manual_seed(seed * 100003 + b)
        init = torch.rand(1, 1, blk, W, device="cuda", generator=g)
        aff = torch.sigmoid(1.5 * torch.randn(1, 9, blk, W, device="cuda", generator=g))
        off = (1.5 * torch.randn(1, 18, blk, W, device="cuda", generator=g)).clamp_(-6, 6)
        lo, hi = max(r0, b * blk) - b * blk, min(r1, (b + 1) * blk) - b * blk
        outs[0].append(init[:, :, lo:hi]); outs[1].append(aff[:, :, lo:hi]); outs[2].append(off[:, :, lo:hi])
    init, aff, off = (torch.cat(o, dim=2).contiguous() for o in outs)
    off[:, 8:10] = 0
    assert init.shape[2] == rows
    return init, aff, off


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--H", type=int, default=32768)
    ap.add_argument("--W", type=int, default=32768)
    ap.add_argument("--T", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = torch.ones(1, 1, 3, 3, device="cuda") * 1.05
    b = torch.full((1,), 0.1, device="cuda")

    # ---- correctness on a small raster: strips == unsharded, bit for bit ----
    Hs, Ws = 1024, 512
    full = make_rows(0, Hs, Ws, 7)
    ref1 = F.spn_forward(*full, w, b, 1, 1.0)
    reff = F.spn_iterate(full[0], full[1] * 0.1, full[2], 3)
    r0, r1, _, _ = strip_bounds(Hs, world, rank, 0)
    band = [t[:, :, r0:r1].contiguous() for t in full]
    sp = StripPropagator(Hs)
    out, status = sp.forward(band[0], band[1], band[2], w, b, 1, 1.0)
    out_hb, status_hb = sp.forward(sp.halo_buffer(band[0], 8), band[1], band[2], w, b, 1, 1.0)
    feats, status2 = sp.iterate(band[0], band[1] * 0.1, band[2], 3)
    ok = torch.equal(out, ref1[:, :, r0:r1]) and all(torch.equal(f, reff[t][:, :, r0:r1]) for t, f in enumerate(feats))
    ok = ok and torch.equal(out_hb, out) and int(status.item()) == 0 and int(status2.item()) == 0 and int(status_hb.item()) == 0
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("strip == unsharded (bitwise, T=1 and T=3):", bool(flag.item()), flush=True)
    del full, ref1, reff, band, out, feats

    # ---- timing on the large raster ----
    H, W, T = args.H, args.W, args.T
    r0, r1, _, _ = strip_bounds(H, world, rank, 0)
    init, aff, off = make_rows(r0, r1, W, 11)
    sp = StripPropagator(H)
    halo = 8  # offsets are clipped to +-6: ceil(6) + 2
    hb = sp.halo_buffer(init, halo)   # the band lives in a buffer with halo room: no copies in the loop

    def step():
        if T == 1:
            return sp.forward(hb, aff, off, w, b, 1, 1.0)[0]
        return sp.iterate(init, aff, off, T, halo=halo, keep_all=False)[0][-1]

    for _ in range(3):
        step()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize(); dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device="cuda", dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        gpix = H * W * T / (ms.item() * 1e-3) / 1e9
        print(json.dumps({"workload": f"row-strip inference {H}x{W} fp32, T={T}, halo {halo} rows", "n_gpus": world,
                          "ms_per_step": ms.item(), "value": gpix, "unit": "Gpix·iter/s",
                          "hbm_gbs_per_gpu": 116 * H * W * T / world / (ms.item() * 1e-3) / 1e9,
                          "halo_bytes_per_exchange_per_rank": 2 * halo * W * 4}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
