"""CPU, gloo, world_size 2: the bucketed gradient reducer reproduces the single-process gradient of
the global batch, skips parameters without gradients, and shard_batch cuts the batch layout."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Toy(nn.Module):
    def __init__(self):
        super().__init__()
        self.a = nn.Linear(6, 5)
        self.b = nn.Linear(5, 3)
        self.unused = nn.Linear(4, 4)       # like Demand's gate.fc: never gets a gradient

    def forward(self, x):
        return self.b(torch.tanh(self.a(x)))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from visuelle2_multimodal_fusion_b200.ddp import GradReducer, shard_batch
    torch.manual_seed(0)
    m = _Toy()
    x = torch.randn(8, 6)
    y = torch.randn(8, 3)
    red = GradReducer(m, bucket_bytes=64)            # tiny buckets -> several buckets
    assert len(red.buckets) > 2
    for it in range(2):                              # twice: the reducer must re-arm
        (xs, ys), _ = shard_batch(((x, y), x), rank, world)
        m.zero_grad(set_to_none=True)
        loss = ((m(xs) - ys) ** 2).mean()
        loss.backward()
        red.finish()
    if rank == 0:
        torch.save({k: (p.grad.clone() if p.grad is not None else None) for k, p in m.named_parameters()}, out)
    # finish() hands the mean over by re-pointing p.grad at its slice of the reduced bucket; reduce_now() -- gradients
    # left in static storage by a CUDA-graph replay -- must write INTO that storage instead
    red.remove()
    red2 = GradReducer(m, bucket_bytes=64, hooks=False)
    m.zero_grad(set_to_none=True)
    (xs, ys), _ = shard_batch(((x, y), x), rank, world)
    ((m(xs) - ys) ** 2).mean().backward()
    static = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    ptrs = {k: g.data_ptr() for k, g in static.items()}
    red2.reduce_now()
    for k, p in m.named_parameters():
        if p.grad is not None:
            assert p.grad is static[k] and p.grad.data_ptr() == ptrs[k], k
    if rank == 0:
        torch.save({k: (p.grad.clone() if p.grad is not None else None) for k, p in m.named_parameters()}, out + ".now")
    dist.destroy_process_group()


def test_bucketed_allreduce_matches_global_batch(tmp_path):
    out = str(tmp_path / "g.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    m = _Toy()
    x = torch.randn(8, 6)
    y = torch.randn(8, 3)
    ((m(x) - y) ** 2).mean().backward()
    now = torch.load(out + ".now")
    for k, p in m.named_parameters():
        if p.grad is None:
            assert got[k] is None and now[k] is None, k
        else:
            assert torch.allclose(got[k], p.grad, atol=1e-6), k
            assert torch.allclose(now[k], p.grad, atol=1e-6), k
