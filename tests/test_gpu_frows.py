"""GPU tests of the SURVEY.md §8f "next" rows built on the hand-written kernels: fused losses (row 1), SWA (row 2),
validation / ensemble path + device-side mAP (row 3), waveform augmentation (row 4).  Checkers are plain torch / numpy /
sklearn restatements of the reference code (cited per test), run on the CPU."""
import copy

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from util import quiet, relerr

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_fused_bce_matches_reference_training_loss():
    """ex_audioset.py:172-192: y_mix = y*lam + y[perm]*(1-lam); BCE-with-logits(reduction none).mean(); gradient."""
    from passt_b200 import loss as PL
    torch.manual_seed(0)
    B, C = 64, 527
    z = (3 * torch.randn(B, C)).requires_grad_(True)
    y = (torch.rand(B, C) < 0.01).float()
    np.random.seed(1)
    perm, lam = PL.draw_mixup(B, 0.3)
    for use_mix in (False, True):
        zc = z.detach().clone().to(DEV).requires_grad_(True)
        got = PL.bce_with_logits(zc, y.to(DEV), perm.to(DEV) if use_mix else None, lam.to(DEV) if use_mix else None)
        (got * 1.7).backward()
        yy = y * lam.reshape(B, 1) + y[perm] * (1.0 - lam.reshape(B, 1)) if use_mix else y
        z.grad = None
        ref = F.binary_cross_entropy_with_logits(z, yy, reduction="none").mean()
        (ref * 1.7).backward()
        assert abs(float(got) - float(ref)) <= 2e-6 * abs(float(ref))
        assert relerr(zc.grad, z.grad) < 1e-5


def test_fused_cross_entropy_mixup_matches_reference():
    """ex_esc50.py:151-169: CE(y_hat, y)*lam + CE(y_hat, y[perm])*(1-lam), mean."""
    from passt_b200 import loss as PL
    torch.manual_seed(1)
    B, C = 32, 50
    z = (2 * torch.randn(B, C)).requires_grad_(True)
    y = torch.randint(C, (B,))
    np.random.seed(2)
    perm, lam = PL.draw_mixup(B, 0.3)
    for use_mix in (False, True):
        zc = z.detach().clone().to(DEV).requires_grad_(True)
        got = PL.cross_entropy(zc, y.to(DEV), perm.to(DEV) if use_mix else None, lam.to(DEV) if use_mix else None)
        got.backward()
        z.grad = None
        if use_mix:
            ref = (F.cross_entropy(z, y, reduction="none") * lam + F.cross_entropy(z, y[perm], reduction="none") * (1. - lam)).mean()
        else:
            ref = F.cross_entropy(z, y, reduction="none").mean()
        ref.backward()
        assert abs(float(got) - float(ref)) <= 2e-6 * abs(float(ref))
        assert relerr(zc.grad, z.grad) < 1e-5


def test_draw_mixup_is_the_reference_draw():
    """helpers/mixup.py:5-12 restated: randperm from torch's CPU generator, beta from numpy's global RNG."""
    from passt_b200 import loss as PL
    torch.manual_seed(5); np.random.seed(6)
    perm, lam = PL.draw_mixup(16, 0.3)
    torch.manual_seed(5); np.random.seed(6)
    rn = torch.randperm(16)
    lambd = np.random.beta(0.3, 0.3, 16).astype(np.float32)
    lambd = np.concatenate([lambd[:, None], 1 - lambd[:, None]], 1).max(1)
    assert torch.equal(perm, rn) and torch.equal(lam, torch.FloatTensor(lambd))


def test_swa_update_matches_reference_formula():
    """helpers/swa_callback.py:246-268: first update copies, then p_swa += (p - p_swa)/(n+1)."""
    from passt_b200.passt import get_model, lighten_model
    from passt_b200.swa import SWAAverager
    with quiet():
        net = lighten_model(get_model(arch="passt_s_swa_p16_128_ap476", pretrained=False), cut_depth=10).to(DEV)
    swa = SWAAverager(net)
    ref = [p.detach().clone().cpu() for p in net.parameters()]
    g = torch.Generator().manual_seed(0)
    n_avg = 0
    for it in range(3):
        with torch.no_grad():
            for p in net.parameters():
                p.add_(0.01 * torch.randn(p.shape, generator=g).to(DEV))
        swa.update()
        for i, p in enumerate(net.parameters()):
            pm = p.detach().cpu()
            ref[i] = pm.clone() if n_avg == 0 else ref[i] + (pm - ref[i]) / (n_avg + 1)
        n_avg += 1
    assert swa.n_averaged == 3
    for a, b in zip(swa.net_swa.parameters(), ref):
        assert torch.allclose(a.detach().cpu(), b, rtol=1e-6, atol=1e-7)
    # the averaged net evaluates with its NEW weights (bf16 operand copies re-cast)
    x = torch.randn(2, 1, 128, 1000, device=DEV)
    fresh = copy.deepcopy(swa.net_swa).eval()
    with torch.no_grad():
        a, _ = swa.net_swa.eval()(x)
        b, _ = fresh(x)
    assert torch.equal(a, b)


def test_validation_step_two_nets_share_one_mel():
    """ex_audioset.py:216-245: net and net_swa on the same spectrogram; out = sigmoid(logits), loss = BCE mean."""
    from passt_b200.passt import get_model, lighten_model
    from passt_b200.preprocess import AugmentMelSTFT
    from passt_b200 import evalpath
    with quiet():
        net = lighten_model(get_model(arch="passt_s_swa_p16_128_ap476", pretrained=False), cut_depth=10).to(DEV).eval()
        mel = AugmentMelSTFT(fmin_aug_range=10, fmax_aug_range=2000).to(DEV).eval()
    net2 = copy.deepcopy(net)
    with torch.no_grad():
        for p in net2.parameters():
            p.mul_(1.01)
    torch.manual_seed(0)
    wave = 0.1 * torch.randn(3, 1, 320000, device=DEV)
    y = (torch.rand(3, 527, device=DEV) < 0.02).float()
    calls = []
    orig = mel.forward
    mel.forward = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
    res = evalpath.validation_step(mel, [("", net), ("swa_", net2)], wave, y)
    mel.forward = orig
    assert len(calls) == 1                                   # ONE frontend pass for both nets
    with torch.no_grad():
        spec = mel(wave.reshape(3, -1)).unsqueeze(1)
        for prefix, n in (("", net), ("swa_", net2)):
            logits, _ = n(spec)
            assert torch.allclose(res[prefix + "out"], torch.sigmoid(logits), atol=1e-6)
            ref = F.binary_cross_entropy_with_logits(logits, y)
            assert abs(float(res[prefix + "val_loss"]) - float(ref)) < 1e-5 * abs(float(ref))
    # ensemble = mean of the logits (models/passt.py:1021-1036)
    ens = evalpath.EnsembleRunner([net, net2])
    with torch.no_grad():
        m, _ = ens(spec)
        assert torch.allclose(m, (net(spec)[0] + net2(spec)[0]) / 2, atol=1e-5)
        assert torch.allclose(ens.predict_proba(spec), torch.sigmoid(m), atol=1e-6)


def test_device_average_precision_matches_sklearn():
    """validation_epoch_end (ex_audioset.py:262-266): metrics.average_precision_score(target, out, average=None)."""
    from sklearn import metrics
    from passt_b200.evalpath import MeanAPMeter
    rng = np.random.RandomState(0)
    n, C = 3000, 37
    tgt = (rng.rand(n, C) < 0.03).astype(np.float32)
    tgt[0] = 1.0                                              # every class has a positive
    out = rng.rand(n, C).astype(np.float32)
    out[:, :10] = np.round(out[:, :10] * 20) / 20             # heavy ties in the first classes
    out += 0.3 * tgt                                          # make it informative
    out[:, 10:20] = np.round(out[:, 10:20], 2)
    meter = MeanAPMeter()
    for i in range(0, n, 1000):
        meter.update(torch.from_numpy(out[i:i + 1000]).to(DEV), torch.from_numpy(tgt[i:i + 1000]).to(DEV))
    ap = meter.average_precision().cpu().numpy()
    ref = metrics.average_precision_score(tgt, out, average=None)
    assert np.allclose(ap, ref, rtol=2e-5, atol=1e-6), np.abs(ap - ref).max()
    assert abs(float(meter.mean_ap()) - ref.mean()) < 1e-5


def test_wave_augment_matches_loader_pipeline():
    """gain -> pad_or_truncate -> roll -> MixupDataset (audioset/dataset.py:107-140, 315-339), restated with numpy/torch
    on the CPU per clip; ragged source lengths (shorter, equal and longer than the clip length)."""
    from passt_b200.waveaug import WaveAugment
    Lc = 32000
    torch.manual_seed(3); np.random.seed(4)
    lens = [32000, 20000, 40000, 32000, 5000, 32001]
    raws = [0.1 * torch.randn(n) + 0.01 for n in lens]
    y = (torch.rand(len(lens), 527) < 0.02).float()
    aug = WaveAugment(clip_length=Lc, gain_augment=7, roll_range=50, wavmix_rate=0.7, wavmix_beta=2)
    d = aug.draw(len(lens))
    assert bool((d.mix_idx >= 0).any()) and bool((d.mix_idx < 0).any())
    got, got_y = aug([r.to(DEV) for r in raws], y.to(DEV), d)

    def prep(b):
        w = raws[b].numpy() * float(d.gain[b])
        w = np.concatenate((w, np.zeros(Lc - len(w), dtype=np.float32))) if len(w) <= Lc else w[:Lc]
        return torch.as_tensor(w).reshape(1, -1).roll(int(d.shift[b]), 1)
    for b in range(len(lens)):
        x1 = prep(b)
        if int(d.mix_idx[b]) >= 0:
            l = float(d.mix_lam[b])
            x2 = prep(int(d.mix_idx[b]))
            x1 = x1 - x1.mean(); x2 = x2 - x2.mean()
            x = x1 * l + x2 * (1. - l)
            x = x - x.mean()
            ty = y[b] * l + y[int(d.mix_idx[b])] * (1. - l)
        else:
            x, ty = x1, y[b]
        assert torch.allclose(got[b].cpu(), x.reshape(-1), atol=2e-6), b
        assert torch.allclose(got_y[b].cpu(), ty, atol=1e-6), b
    # fixed-length batch input, no mixup, no targets
    batch = torch.stack([r[:5000] for r in raws]).to(DEV)
    aug2 = WaveAugment(clip_length=8000, gain_augment=0, roll_range=0)
    out, none_y = aug2(batch, None, aug2.draw(len(lens)))
    assert none_y is None and torch.equal(out[:, :5000], batch) and float(out[:, 5000:].abs().max()) == 0.0
