"""Pin the CPU oracle against the reference implementation itself (imported from /root/reference through the shim).
Skipped where the reference tree is absent (the GPU box) — there the committed golden fixture takes over."""
import pytest
import torch

from ref_shim import load_reference, quiet, reference_available
from oracle import passt_oracle as O

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present")


def _ref_mel():
    _, rpre = load_reference()
    with quiet():
        return rpre.AugmentMelSTFT(n_mels=128, sr=32000, win_length=800, hopsize=320, n_fft=1024, freqm=48, timem=192,
                                   htk=False, fmin=0.0, fmax=None, norm=1, fmin_aug_range=10, fmax_aug_range=2000)


@pytest.mark.parametrize("training", [False, True])
@pytest.mark.parametrize("L", [48000, 33001])
def test_mel_bit_exact(training, L):
    mel = _ref_mel().train(training)
    cfg = O.MelCfg()
    torch.manual_seed(0)
    wave = 0.1 * torch.randn(2, L)
    torch.manual_seed(3)
    with quiet():
        ref = mel(wave)
    torch.manual_seed(3)
    d = O.draw_mel(cfg, training, 2)
    mine = O.mel_frontend(wave, cfg, d, training)
    assert torch.equal(ref, mine)


def test_mel_banks_match_torchaudio():
    import torchaudio
    for fmin, fmax in [(0.0, 15000.0), (7.0, 14321.0), (9.0, 16000.0)]:
        ref, _ = torchaudio.compliance.kaldi.get_mel_banks(128, 1024, 32000, fmin, fmax, 100.0, -500.0, 1.0)
        assert torch.equal(ref, O.kaldi_mel_banks(128, 1024, 32000, fmin, fmax))


@pytest.mark.parametrize("kw,T", [(dict(s_patchout_t=40, s_patchout_f=4), 1000), (dict(u_patchout=400), 1000),
                                  (dict(s_patchout_t=10, s_patchout_f=3, n_classes=50), 500)])
def test_net_train_forward_backward_light(kw, T):
    """3-block model (reference lighten_model cut_depth=9): logits, features, draws and gradients."""
    rp, _ = load_reference()
    cfg12 = O.NetCfg(**kw)
    with quiet():
        net = rp.get_model(arch="passt_s_swa_p16_128_ap476", pretrained=False, n_classes=cfg12.n_classes,
                           u_patchout=cfg12.u_patchout, s_patchout_t=cfg12.s_patchout_t,
                           s_patchout_f=cfg12.s_patchout_f)
        net.load_state_dict(O.synth_params(cfg12, 2), strict=True)
        net = rp.lighten_model(net, cut_depth=9)        # keeps blocks 0, 10, 11
    cfg = O.NetCfg(depth=3, **kw)
    p = {}
    remap = {0: 0, 10: 1, 11: 2}
    for k, v in O.synth_params(cfg12, 2).items():
        if k.startswith("blocks."):
            i = int(k.split(".")[1])
            if i in remap:
                p[k.replace(f"blocks.{i}.", f"blocks.{remap[i]}.", 1)] = v
        else:
            p[k] = v
    p = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    torch.manual_seed(1)
    x = torch.randn(1, 1, 128, T)
    net.train()
    torch.manual_seed(8)
    with quiet():
        ref_logits, ref_feat = net(x)
    torch.manual_seed(8)
    d = O.draw_patchout(cfg, 12, (T - 16) // 10 + 1, True)
    lg, ft = O.passt_forward(p, x, cfg, d)
    assert torch.equal(ref_logits, lg) and torch.equal(ref_feat, ft)
    ref_logits.sum().backward()
    lg.sum().backward()
    ref_grads = dict(net.named_parameters())
    for k, v in p.items():
        if k.startswith("head_dist"):
            assert v.grad is None and ref_grads[k].grad is None
            continue
        assert torch.allclose(v.grad, ref_grads[k].grad, rtol=1e-5, atol=1e-7), k
