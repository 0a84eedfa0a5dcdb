"""Where does the fused halo exchange cost time?  torchrun --nproc-per-node 2 tools/strip_peer_probe.py
Times, on 4096 x 32768 rows per rank: the plain strip kernel on torch buffers, the plain kernel reading the ring's
(peer-mapped) buffer, and the PEER kernel with parts of the exchange switched off (JSPSR_STRIP_PEER_DEBUG, timing only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from jspsr_b200 import functional as F
from jspsr_b200.strips import StripPropagator
import bench

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
rows, W, halo = 4096, 32768, 8
H_img = rows * world
init, aff, off = bench.strip_rows(torch, dev, rank * rows, (rank + 1) * rows, W, 11)
aff = aff * 0.1
w = torch.full((1, 1, 3, 3), 1.05, device=dev); b = torch.full((1,), 0.1, device=dev)
status = torch.zeros(1, dtype=torch.int32, device=dev)
out = torch.empty(1, 1, rows, W, device=dev)
sp = StripPropagator(H_img, rank, world)
ring = sp.peer_ring(rows, W, halo, n_buf=2)
ring.load(init)


def timed(fn, n=8):
    for _ in range(2):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def say(name, ms):
    if rank == 0:
        print(f"{name:70s} {ms:.3f} ms", flush=True)

# the plain kernel sees its own band only: the strip is treated as a whole image of `rows` rows
say("plain kernel, torch buffers (no exchange, band as an image)",
    timed(lambda: F.spn_forward_strip(init, aff, off, w, b, 1, 1.0, rows, 0, 0, status, out=out)))
buf = ring.buf(ring.cur)
say("plain kernel, DEM read from the ring buffer (peer-mapped memory)",
    timed(lambda: F.spn_forward_strip(buf[:, :, ring.top:ring.top + rows], aff, off, w, b, 1, 1.0, rows, 0, 0, status, out=out)))
say("plain kernel, out written into the ring's other buffer",
    timed(lambda: F.spn_forward_strip(init, aff, off, w, b, 1, 1.0, rows, 0, 0, status, out=ring.interior(1 - ring.cur))))
for dbg, what in ((0, "full protocol"), (1, "plain tile order"), (2, "no row copy"), (4 | 8, "no waits, no push"),
                  (1 | 4 | 8, "no waits, no push, plain order"), (4, "no waits"), (0, "full protocol again")):
    os.environ["JSPSR_STRIP_PEER_DEBUG"] = str(dbg)
    say(f"PEER T = 1 (push kernel + forward), debug {dbg}: {what}", timed(lambda: sp.forward_peer(ring, aff, off, w, b, 1, 1.0, out=out)))
    say(f"PEER T = 6 per application, debug {dbg}: {what}", timed(lambda: sp.iterate_peer(ring, aff, off, 6), n=3) / 6)
os.environ.pop("JSPSR_STRIP_PEER_DEBUG")
ring.close()
if world > 1:
    dist.destroy_process_group()
