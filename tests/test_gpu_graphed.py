"""CUDA-graph replay of the train step == the eager train step (same kernels, same draws), step by step."""
import copy

import pytest
import torch
import torch.nn.functional as F

from util import quiet

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _make(seed):
    from passt_b200.passt import get_model, lighten_model
    from passt_b200.preprocess import AugmentMelSTFT
    torch.manual_seed(seed)
    with quiet():
        net = get_model(arch="passt_s_swa_p16_128_ap476", pretrained=False, s_patchout_t=40, s_patchout_f=4)
        net = lighten_model(net, cut_depth=9).to(DEV).train()          # 3 blocks
        mel = AugmentMelSTFT(freqm=0, timem=0, fmin_aug_range=10, fmax_aug_range=2000).to(DEV).train()
    return mel, net


def test_graphed_step_matches_eager():
    from passt_b200.graphed import GraphedTrainStep
    B = 4
    torch.manual_seed(0)
    waves = [0.1 * torch.randn(B, 320000, device=DEV) for _ in range(4)]
    y = (torch.rand(B, 527, device=DEV) < 0.05).float()
    mel_a, net_a = _make(1)
    mel_b, net_b = copy.deepcopy(mel_a), copy.deepcopy(net_a)
    params = lambda n: [p for k, p in n.named_parameters() if not k.startswith("head_dist")]
    opt_a = torch.optim.AdamW(params(net_a), lr=1e-3, weight_decay=1e-4, fused=True)
    opt_b = torch.optim.AdamW(params(net_b), lr=1e-3, weight_decay=1e-4, fused=True, capturable=True)
    loss_fn = F.binary_cross_entropy_with_logits
    # graphed: construction runs warm-up steps with lr > 0, so rebuild identical starting weights afterwards
    sd0 = copy.deepcopy(net_a.state_dict())
    step = GraphedTrainStep(mel_b, net_b, opt_b, loss_fn, waves[0], y, warmup=2)
    net_b.load_state_dict(sd0)
    for st in opt_b.state.values():                 # reset AdamW moments / step counters accumulated in warm-up
        for k, v in st.items():
            if torch.is_tensor(v):
                v.zero_()
    losses_a, losses_b = [], []
    for i, w in enumerate(waves):
        torch.manual_seed(100 + i)                  # same host draws for step i in both paths
        spec = mel_a(w).unsqueeze(1)
        logits, _ = net_a(spec)
        la = loss_fn(logits, y)
        opt_a.zero_grad(set_to_none=True)
        la.backward()
        opt_a.step()
        losses_a.append(float(la))
        torch.manual_seed(100 + i)
        lb = step(w, y)
        losses_b.append(float(lb))
    # step 0 is bit-identical; later steps differ only through the summation order of the fp32 atomics / TMA
    # reduce-adds in the gradient kernels, which AdamW's sign-like update amplifies slightly
    assert losses_a[0] == losses_b[0]
    assert losses_a == pytest.approx(losses_b, rel=2e-3), (losses_a, losses_b)
    # AdamW's update is ~lr*sign(g) in the first steps, so elements whose gradient is ~0 may move in opposite
    # directions under a different fp32 summation order: bound = 2*lr*steps per element, and the bulk must agree
    lr, nsteps = 1e-3, len(waves)
    for (ka, pa), (kb, pb) in zip(net_a.named_parameters(), net_b.named_parameters()):
        diff = (pa - pb).abs()
        assert diff.max().item() <= 2 * lr * nsteps * 1.05, (ka, diff.max().item())
        assert diff.mean().item() < 0.25 * lr, (ka, diff.mean().item())
    # eager inference after graph training sees the updated weights (bf16 cache invalidated)
    net_a.eval(); net_b.eval()
    with torch.no_grad():
        xa = mel_a.eval()(waves[0]).unsqueeze(1)
        la, lb = net_a(xa)[0], net_b(xa)[0]
        assert ((la - lb).abs().max() / la.abs().max()).item() < 2e-2


def test_graphed_inference_matches_eager_and_resyncs():
    """GraphedInference (the cfg1 / cfg4 bench path) == the eager eval forward, also after the weights were edited and
    resync() re-cast the bf16 operand copies."""
    from passt_b200.graphed import GraphedInference
    B = 3
    mel, net = _make(3)
    mel.eval(); net.eval()
    torch.manual_seed(1)
    waves = [0.1 * torch.randn(B, 320000, device=DEV) for _ in range(3)]

    def eager(w):
        with torch.no_grad():
            return net(mel(w).unsqueeze(1))[0].clone()

    g = GraphedInference(mel, net, waves[0])
    for w in waves:
        assert torch.equal(g(w), eager(w))
    with torch.no_grad():
        for p in net.parameters():
            p.mul_(1.02)
    ref = eager(waves[1])                 # the eager forward notices the in-place edit (version counter) and re-casts
    g.resync()
    got = g(waves[1]).clone()
    assert torch.equal(got, ref)
    assert not torch.equal(got, g(waves[2]))
