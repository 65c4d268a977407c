"""GPU smoke/parity tests at the other BASELINE.json configurations' shapes (beyond cfg1/cfg2 in test_gpu_parity.py):
cfg3 (u_patchout + fused spectrogram mixup), cfg4 (passt_l depth 7 inference), cfg5 (ESC-50 head, 5 s clips) and the
size-independent properties of the path at full size (finite outputs, gradient of every used parameter present,
determinism for a fixed seed)."""
import pytest
import torch

from util import build_net, depth2_params, quiet, relerr

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _oracle():
    from oracle import passt_oracle as O
    return O


def test_cfg4_passt_l_eval_parity_small_batch():
    O = _oracle()
    cfg = O.NetCfg(depth=7)
    params = O.synth_params(cfg, seed=5)
    net = build_net(cfg, params, DEV, arch="passt_l_kd_p16_128_ap47").eval()
    torch.manual_seed(1)
    x = torch.randn(2, 1, 128, 1000)
    with torch.no_grad():
        logits, feats = net(x.to(DEV))
        ref_logits, ref_feats = O.passt_forward(params, x, cfg, O.StepDraws())
    assert relerr(logits, ref_logits) < 1e-2 and relerr(feats, ref_feats) < 1e-2


def test_cfg4_full_batch_runs_and_is_deterministic():
    """passt_l, no patchout, N = 1190, batch 256 (304 640 tokens): same input twice -> identical logits; finite."""
    from passt_b200.passt import get_model
    with quiet():
        net = get_model(arch="passt_l_kd_p16_128_ap47", pretrained=False).to(DEV).eval()
    torch.manual_seed(0)
    x = torch.randn(256, 1, 128, 1000, device=DEV)
    with torch.no_grad():
        a, _ = net(x)
        b, _ = net(x)
    assert torch.isfinite(a).all() and torch.equal(a, b)
    # batch independence: clip i of the big batch equals the same clip run alone (same kernels, different grid sizes)
    with torch.no_grad():
        c, _ = net(x[5:7])
    assert relerr(a[5:7], c) < 2e-3


def test_cfg3_fused_mixup_matches_explicit_mixup():
    """x*lam + x[perm]*(1-lam) folded into the patch gather == mixing the spectrograms first (ex_audioset.py:173-177)."""
    O = _oracle()
    kw = dict(u_patchout=400)
    cfg12 = O.NetCfg(**kw)
    params = O.synth_params(cfg12, seed=7)
    net = build_net(cfg12, params, DEV, cut_depth=10).train()
    B = 4
    torch.manual_seed(3)
    x = torch.randn(B, 1, 128, 1000, device=DEV)
    perm = torch.randperm(B).to(DEV)
    lam = torch.rand(B, device=DEV) * 0.5 + 0.5
    mixed = x * lam.view(B, 1, 1, 1) + x[perm] * (1 - lam.view(B, 1, 1, 1))
    torch.manual_seed(11)
    ref, _ = net(mixed)
    net.fused_mixup(perm, lam)
    torch.manual_seed(11)
    got, _ = net(x)
    assert relerr(got, ref) < 5e-3      # bf16 rounding of (mixed patch) vs same value computed from fp32 inputs


def test_cfg5_esc50_train_step_all_grads_present():
    from passt_b200.passt import get_model
    with quiet():
        net = get_model(arch="passt_s_kd_p16_128_ap486", pretrained=False, n_classes=50, s_patchout_t=10,
                        s_patchout_f=3).to(DEV).train()
    torch.manual_seed(0)
    x = torch.randn(16, 1, 128, 500, device=DEV)          # 5 s clips -> 500 frames
    y = torch.randint(50, (16,), device=DEV)
    logits, _ = net(x)
    assert logits.shape == (16, 50) and net.last_plan.ntok == 353
    torch.nn.functional.cross_entropy(logits, y).backward()
    for n, p in net.named_parameters():
        if n.startswith("head_dist"):
            assert p.grad is None
        else:
            assert p.grad is not None and torch.isfinite(p.grad).all(), n
    # time-embedding gradient is non-zero only on the kept columns (offset + kept patch columns)
    plan = net.last_plan
    g = net.time_new_pos_embed.grad[0, :, 0, :].abs().sum(0)
    kept = torch.zeros_like(g, dtype=torch.bool)
    kept[(plan.toffset + plan.t_keep).to(DEV)] = True
    assert (g[~kept] == 0).all() and (g[kept] > 0).all()


def test_long_clip_20s_variant():
    """passt_s_f128_20sec: input_tdim 2000, 20 s clip -> 199 patch columns < 200 embedding columns (random offset)."""
    from passt_b200.passt import get_model
    with quiet():
        net = get_model(arch="passt_s_f128_20sec_p16_s10_ap474", pretrained=False, input_tdim=2000).to(DEV).eval()
    x = torch.randn(1, 1, 128, 2000, device=DEV)
    with torch.no_grad():
        logits, _ = net(x)
    assert logits.shape == (1, 527) and torch.isfinite(logits).all() and net.last_plan.ntok == 12 * 199 + 2


def test_bf16_weight_cache_tracks_fused_optimizer_updates():
    """torch.optim.AdamW(fused=True) updates parameters without bumping Tensor._version: the tensor-core weight copies
    must still follow the fp32 masters (every training forward re-casts; the first eval forward afterwards too)."""
    import copy
    from passt_b200.passt import get_model, lighten_model
    with quiet():
        net = lighten_model(get_model(arch="passt_s_swa_p16_128_ap476", pretrained=False, s_patchout_t=40,
                                      s_patchout_f=4), cut_depth=10).to(DEV).train()
    opt = torch.optim.AdamW([p for n, p in net.named_parameters() if not n.startswith("head_dist")], lr=1e-2, fused=True)
    torch.manual_seed(0)
    x = torch.randn(4, 1, 128, 1000, device=DEV)
    y = (torch.rand(4, 527, device=DEV) < 0.05).float()
    for _ in range(3):
        logits, _ = net(x)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
    net.eval()
    fresh = copy.deepcopy(net)            # deepcopy starts with an empty weight cache -> casts from the fp32 masters
    with torch.no_grad():
        a, _ = net(x)
        b, _ = fresh(x)
    assert torch.equal(a, b)


def test_stride16_arch_eval_parity():
    """passt_s_p16_s16_128_ap468 as fine-tuned on FSD50K (fsd50k/README.md:47-77): no patch overlap, grid 8 x 62,
    200 classes; depth cut to 2 so the CPU oracle stays quick."""
    O = _oracle()
    cfg = O.NetCfg(fstride=16, tstride=16, n_classes=200)
    params = O.synth_params(cfg, seed=11)
    net = build_net(cfg, params, DEV, arch="passt_s_p16_s16_128_ap468", cut_depth=10).eval()
    cfg2 = O.NetCfg(fstride=16, tstride=16, n_classes=200, depth=2)
    torch.manual_seed(2)
    x = torch.randn(2, 1, 128, 1000)
    with torch.no_grad():
        logits, feats = net(x.to(DEV))
        ref_logits, ref_feats = O.passt_forward(depth2_params(params), x, cfg2, O.StepDraws())
    assert logits.shape == (2, 200) and net.last_plan.ntok == 8 * 62 + 2
    assert relerr(logits, ref_logits) < 1e-2 and relerr(feats, ref_feats) < 1e-2


def test_waveform_to_logits_wrapper_matches_oracle_chain():
    """get_basic_model(mode="logits") on raw audio (README.md:49-64 usage: wave[B, samples] @ 32 kHz -> logits):
    frontend kernel + network against the oracle's mel_frontend + passt_forward, eval mode."""
    O = _oracle()
    from passt_b200.wrapper import get_basic_model
    cfg = O.NetCfg()
    params = O.synth_params(cfg, seed=13)
    with quiet():
        model = get_basic_model(mode="logits", pretrained=False)
    model.net.load_state_dict(params, strict=True)
    from passt_b200 import passt as P
    model.net = P.lighten_model(model.net, cut_depth=10)
    model = model.to(DEV).eval()
    torch.manual_seed(4)
    wave = torch.randn(2, 320000) * 0.1
    with torch.no_grad():
        logits = model(wave.to(DEV))
        mcfg = O.MelCfg()
        spec = O.mel_frontend(wave, mcfg, O.StepDraws(fmin=0.0, fmax=mcfg.resolved_fmax()), training=False)
        ref_logits, _ = O.passt_forward(depth2_params(params), spec.unsqueeze(1), O.NetCfg(depth=2), O.StepDraws())
    assert logits.shape == (2, 527)
    assert relerr(logits, ref_logits) < 1e-2


def test_empty_batch_returns_empty_outputs():
    """B = 0 goes through like in the reference (empty tensors), consuming the same random draws."""
    from passt_b200.passt import get_model
    from passt_b200.preprocess import AugmentMelSTFT
    with quiet():
        net = get_model(arch="passt_s_swa_p16_128_ap476", pretrained=False, s_patchout_t=40, s_patchout_f=4).to(DEV).train()
        mel = AugmentMelSTFT(fmin_aug_range=10, fmax_aug_range=2000).to(DEV).train()
    torch.manual_seed(3)
    spec = mel(torch.zeros(0, 320000, device=DEV))
    assert spec.shape == (0, 128, 1000)
    logits, feats = net(spec.unsqueeze(1))
    assert logits.shape == (0, 527) and feats.shape == (0, 768)
    after_empty = torch.randint(1 << 30, (1,)).item()
    torch.manual_seed(3)
    spec = mel(torch.zeros(2, 320000, device=DEV))
    net(spec.unsqueeze(1))
    assert torch.randint(1 << 30, (1,)).item() == after_empty      # identical CPU-generator consumption
