"""Batch-sharded training of the propagation layer under torch DistributedDataParallel (NCCL, 2 GPUs):
DDP gradients of `w`, `b` and of an upstream conv equal the single-process gradients on the concatenated batch
(the hot path itself has no cross-sample coupling).  Skipped on hosts with fewer than 2 GPUs."""
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


class TinyHead(torch.nn.Module):
    """A stand-in for Generator + PostProcessor: 1x1 convs make (weight, offset) from a feature map."""

    def __init__(self):
        super().__init__()
        import jspsr_b200
        self.conv_weight = torch.nn.Conv2d(8, 9, 1)
        self.conv_offset = torch.nn.Conv2d(8, 18, 1)
        self.postprocessor = jspsr_b200.PostProcessor(3, True, 1.0)

    def forward(self, dem, feat):
        weight = torch.sigmoid(self.conv_weight(feat))
        offset = self.conv_offset(feat)
        return self.postprocessor(dem.detach(), weight, offset)


def _data(n, device):
    g = torch.Generator(device="cpu").manual_seed(11)
    dem = torch.rand(n, 1, 64, 128, generator=g).to(device)
    feat = torch.randn(n, 8, 64, 128, generator=g).to(device)
    gt = torch.rand(n, 1, 64, 128, generator=g).to(device)
    return dem, feat, gt


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.manual_seed(0)
        model = TinyHead().cuda()
        ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[rank])
        dem, feat, gt = _data(8, f"cuda:{rank}")
        sl = slice(rank * 4, (rank + 1) * 4)
        loss = (ddp(dem[sl], feat[sl]) - gt[sl]).square().mean()
        loss.backward()
        grads = {n: p.grad.detach().cpu() for n, p in model.named_parameters()}
        if rank == 0:
            torch.manual_seed(0)
            ref = TinyHead().cuda()
            ref.load_state_dict(model.state_dict())
            (ref(dem, feat) - gt).square().mean().backward()
            out = {}
            for n, p in ref.named_parameters():
                g_ref = p.grad.detach().cpu()
                out[n] = float((grads[n] - g_ref).abs().max() / g_ref.abs().max().clamp_min(1e-12))
            q.put(out)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_ddp_gradients_match_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert set(res) == {"conv_weight.weight", "conv_weight.bias", "conv_offset.weight", "conv_offset.bias",
                        "postprocessor.w", "postprocessor.b"}
    for name, rel in res.items():
        assert rel < 2e-5, (name, rel)   # fp32 reduction order differs between the two groupings
