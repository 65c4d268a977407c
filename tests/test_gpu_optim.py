"""FusedAdamW (passt_b200/optim.py, kernel passt_b200/csrc/optim.cu) against torch.optim.AdamW — the optimizer the
reference's get_optimizer builds (ex_audioset.py:104-109) — on the GPU."""
import pytest
import torch

from util import build_net, quiet, relerr

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_fused_adamw_matches_torch_adamw_over_steps():
    from passt_b200.optim import FusedAdamW
    torch.manual_seed(0)
    shapes = [(768, 768), (527, 768), (527,), (1, 1, 768), (3, 5, 7), (4096 * 3 + 5,)]
    pa = [torch.nn.Parameter(torch.randn(s, device=DEV)) for s in shapes]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    kw = dict(lr=3e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.05)
    oa = FusedAdamW(pa, **kw)
    ob = torch.optim.AdamW(pb, **kw)
    for step in range(5):
        if step == 3:                       # an LR scheduler changing the group entry is honoured
            for o in (oa, ob):
                o.param_groups[0]["lr"] = 1e-3
        for a, b in zip(pa, pb):
            g = torch.randn_like(a)
            a.grad = g.clone()
            b.grad = g.clone()
        oa.step()
        ob.step()
    for a, b in zip(pa, pb):
        assert relerr(a, b) < 2e-6
        assert relerr(oa.state[a]["exp_avg"], ob.state[b]["exp_avg"]) < 2e-6
        assert relerr(oa.state[a]["exp_avg_sq"], ob.state[b]["exp_avg_sq"]) < 2e-6
    assert float(oa.state[pa[0]]["step"]) == 5.0
    # state_dict round trip into a fresh optimizer continues identically
    pc = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    oc = FusedAdamW(pc, **kw)
    oc.load_state_dict(oa.state_dict())
    for a, c in zip(pa, pc):
        g = torch.randn_like(a)
        a.grad = g.clone()
        c.grad = g.clone()
    oa.step()
    oc.step()
    for a, c in zip(pa, pc):
        assert torch.equal(a, c)


def test_fused_adamw_refreshes_bf16_weight_copies_of_the_network():
    """attach(net): after a step the cached bf16 operand copies equal the updated fp32 weights, the next training
    forward skips its own refresh, and the result matches the same step taken with torch.optim.AdamW."""
    from oracle import passt_oracle as O
    from passt_b200.optim import FusedAdamW
    cfg = O.NetCfg(s_patchout_t=40, s_patchout_f=4)
    params = O.synth_params(cfg, seed=21)
    nets, opts = [], []
    for own in (True, False):
        net = build_net(cfg, params, DEV, cut_depth=10).train()
        ps = [p for n, p in net.named_parameters() if not n.startswith("head_dist")]
        opt = FusedAdamW(ps, lr=1e-3, weight_decay=1e-2).attach(net) if own else \
            torch.optim.AdamW(ps, lr=1e-3, weight_decay=1e-2)
        nets.append(net); opts.append(opt)
    torch.manual_seed(5)
    x = torch.randn(4, 1, 128, 1000, device=DEV)
    y = (torch.rand(4, 527, device=DEV) < 0.01).float()
    losses = []
    for net, opt in zip(nets, opts):
        torch.manual_seed(9)                       # same patchout draws for both
        for _ in range(3):
            opt.zero_grad(set_to_none=True)
            logits, _ = net(x)
            loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, y)
            loss.backward()
            opt.step()
        losses.append(float(loss))
    assert abs(losses[0] - losses[1]) < 2e-3 * abs(losses[1])
    net = nets[0]
    assert net._wcache.fresh_from_optimizer
    w = net.blocks[0].mlp.fc1.weight
    wb = net._wcache._store[(w.data_ptr(), tuple(w.shape))][1]
    assert torch.equal(wb, w.detach().bfloat16())
    for (n0, p0), (n1, p1) in zip(nets[0].named_parameters(), nets[1].named_parameters()):
        if not n0.startswith("head_dist"):
            assert (p0 - p1).abs().max() < 3 * 3 * 1e-3 + 1e-6, n0     # <= lr per step per element (Adam sign flips)
