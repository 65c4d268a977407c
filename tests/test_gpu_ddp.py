"""2-GPU data-parallel gradient equivalence (needs >= 2 CUDA devices: `gpurun --gpus 2 -- python -m pytest
tests/test_gpu_ddp.py -m gpu`; skipped on a 1-GPU box).

Two NCCL ranks x B/2 clips with passt_b200.ddp.GradAllReducer (chunked all-reduce overlapped with the hand-written
backward) must reproduce the gradient of ONE rank on the global batch of B clips: same weights, same patchout draws
(same CPU seed on every rank), loss = mean BCE over the local batch, gradients averaged over ranks."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_net(dev, depth_cut):
    from util import quiet
    from passt_b200.passt import get_model, lighten_model
    torch.manual_seed(0)
    with quiet():
        net = get_model(arch="passt_s_swa_p16_128_ap476", pretrained=False, s_patchout_t=40, s_patchout_f=4)
        if depth_cut:
            net = lighten_model(net, cut_depth=depth_cut)
    # perturb the structurally-zero parameters so that every gradient path is exercised
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for p in net.parameters():
            p.add_(0.01 * torch.randn(p.shape, generator=g))
    return net.to(dev).train()


def _batch(B):
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, 1, 128, 1000, generator=g)
    y = (torch.rand(B, 527, generator=g) < 0.05).float()
    return x, y


def _worker(rank, world, port, B, depth_cut, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch.distributed as dist
    import torch.nn.functional as F
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from passt_b200.ddp import GradAllReducer
    net = _make_net(dev, depth_cut)
    red = GradAllReducer(net, min_chunk_elems=1 << 20)
    x, y = _batch(B)
    lo, hi = rank * B // world, (rank + 1) * B // world
    torch.manual_seed(77)                       # identical patchout draws on every rank
    logits, _ = net(x[lo:hi].to(dev))
    loss = F.binary_cross_entropy_with_logits(logits, y[lo:hi].to(dev))
    loss.backward()
    red.all_reduce()
    red.check_grads_alias(net)
    torch.cuda.synchronize()
    out[rank] = {k: p.grad.detach().cpu() for k, p in net.named_parameters() if p.grad is not None}
    dist.destroy_process_group()


@pytest.mark.parametrize("depth_cut", [9])
def test_two_ranks_reproduce_the_global_batch_gradient(depth_cut):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    import torch.multiprocessing as mp
    import torch.nn.functional as F
    from util import grad_metrics
    B, world = 8, 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), B, depth_cut, out), nprocs=world, join=True)
    # single rank, global batch
    dev = torch.device("cuda", 0)
    net = _make_net(dev, depth_cut)
    x, y = _batch(B)
    torch.manual_seed(77)
    logits, _ = net(x.to(dev))
    F.binary_cross_entropy_with_logits(logits, y.to(dev)).backward()
    ref = {k: p.grad.detach().cpu() for k, p in net.named_parameters() if p.grad is not None}
    assert set(ref) == set(out[0]) == set(out[1])
    worst = worst_l2 = 0.0
    for k, g in ref.items():
        assert torch.equal(out[0][k], out[1][k]), k          # the all-reduce leaves bit-identical replicas
        m = grad_metrics(out[0][k], g)
        worst = max(worst, m["relmax"])
        # not bit-equal: every gradient is a long fp32 sum over tokens with heavy cancellation, and its order differs
        # (split-K factors of the weight-gradient GEMMs depend on the local token count, TMA reduce-adds / atomics are
        # unordered, NCCL adds the two halves last): measured worst 2.2e-4 in the max-norm, 2e-5 in relative L2
        worst_l2 = max(worst_l2, m["rel_l2"])
        assert m["relmax"] < 1e-3 and m["rel_l2"] < 2e-4 and m["cos"] > 0.999999, (k, m)
    print(f"2-rank vs 1-rank gradient: worst relmax {worst:.2e}, worst rel_l2 {worst_l2:.2e}")
