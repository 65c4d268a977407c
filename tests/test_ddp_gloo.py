"""world_size-2 gloo test of the gradient exchange logic (chunked, coalesced, averaged) on CPU."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from passt_b200.ddp import GradAllReducer

    class Dummy(torch.nn.Module):
        pass

    net = Dummy()
    red = GradAllReducer(net, min_chunk_elems=1000)
    assert net._grad_chunk_hook == red.on_chunk_ready
    n = 10_000
    flat = torch.arange(n, dtype=torch.float32) * (rank + 1)
    # the backward pass reports chunks from the tail towards the head, some smaller than min_chunk (coalesced)
    bounds = [(9000, 10000), (8700, 9000), (8000, 8700), (3000, 8000), (2990, 3000), (0, 2990)]
    for lo, hi in bounds:
        net._grad_chunk_hook(flat, lo, hi)
    red.all_reduce()
    want = torch.arange(n, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
    ok = torch.allclose(flat, want)
    # a second step must work with a fresh buffer (no stale state)
    flat2 = torch.ones(n) * (rank + 1)
    net._grad_chunk_hook(flat2, 0, n)
    red.finish()
    ok = ok and torch.allclose(flat2, torch.full((n,), sum(range(1, world + 1)) / world))
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_chunked_allreduce_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))
