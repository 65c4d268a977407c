"""Drop-in boundary under the reference's DEFAULT configuration (ex_audioset.py:74,79: precision=16, compile=True):
torch.compile(net), fp16 autocast + GradScaler, and a ba3l-style Ingredient that captures default keyword arguments."""
import inspect

import pytest
import torch
import torch.nn.functional as F

from util import quiet, relerr

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _small_net():
    from passt_b200.passt import get_model, lighten_model
    torch.manual_seed(0)
    with quiet():
        net = lighten_model(get_model(arch="passt_s_swa_p16_128_ap476", pretrained=False, s_patchout_t=40,
                                      s_patchout_f=4), cut_depth=10)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for p in net.parameters():
            p.add_(0.01 * torch.randn(p.shape, generator=g))
    return net.to(DEV).train()


def test_torch_compile_of_the_net_is_an_opaque_call_with_gradients():
    """self.net = torch.compile(self.net) (ex_audioset.py:132-135): compiled forward + backward == eager engine."""
    net = _small_net()
    cnet = torch.compile(net)
    torch.manual_seed(3)
    x = torch.randn(4, 1, 128, 1000, device=DEV)
    y = (torch.rand(4, 527, device=DEV) < 0.05).float()
    torch.manual_seed(9)
    la, _ = net(x)
    F.binary_cross_entropy_with_logits(la, y).backward()
    ga = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
    net.zero_grad(set_to_none=True)
    torch.manual_seed(9)
    lb, emb = cnet(x)
    F.binary_cross_entropy_with_logits(lb, y).backward()
    assert torch.equal(la, lb) and emb.shape == (4, 768)
    for k, p in net.named_parameters():
        if k in ga:
            assert relerr(p.grad, ga[k]) < 1e-4, k       # same kernels; fp32 atomics order may differ
    # the compiled wrapper exposes the parameters the way the reference's optimizer construction expects
    assert sum(p.numel() for p in cnet.parameters()) == sum(p.numel() for p in net.parameters())
    net.eval()
    with torch.no_grad():
        a, _ = net(x)
        b, _ = cnet(x)
    assert torch.equal(a, b)


def test_fp16_autocast_and_gradscaler_step():
    """PL precision=16 (ex_audioset.py:74; model_speed_test :399-421): torch.autocast(fp16) around the step and a
    GradScaler on the loss.  The engine keeps its own arithmetic (bf16 operands / fp32 accumulate, fp32 logits), so the
    scaled backward must be finite and equal to the unscaled gradients after unscale_."""
    from passt_b200.preprocess import AugmentMelSTFT
    net = _small_net()
    with quiet():
        mel = AugmentMelSTFT(freqm=0, timem=0, fmin_aug_range=1, fmax_aug_range=1).to(DEV).train()
    torch.manual_seed(4)
    wave = 0.1 * torch.randn(4, 1, 320000, device=DEV)
    y = (torch.rand(4, 527, device=DEV) < 0.05).float()
    opt = torch.optim.AdamW([p for n, p in net.named_parameters() if not n.startswith("head_dist")], lr=1e-4)
    # reference: unscaled fp32-boundary run
    torch.manual_seed(5)
    with torch.no_grad():
        spec = mel(wave.reshape(4, -1)).unsqueeze(1)
    logits, _ = net(spec)
    F.binary_cross_entropy_with_logits(logits, y).backward()
    ref = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
    net.zero_grad(set_to_none=True)
    scaler = torch.amp.GradScaler("cuda")
    torch.manual_seed(5)
    with torch.autocast("cuda", dtype=torch.float16):
        with torch.no_grad():
            spec16 = mel(wave.reshape(4, -1)).unsqueeze(1)
        y_hat, embed = net(spec16)
        loss = F.binary_cross_entropy_with_logits(y_hat, y, reduction="none").mean()
    assert y_hat.dtype == torch.float32 and torch.equal(y_hat, logits)
    scaler.scale(loss).backward()
    scaler.unscale_(opt)
    for k, p in net.named_parameters():
        if k in ref:
            assert torch.isfinite(p.grad).all(), k
            assert relerr(p.grad, ref[k]) < 5e-3, k       # bf16 rounding of the 65536x-scaled backward operands
    before = net.head[1].weight.detach().clone()
    scaler.step(opt)
    scaler.update()
    assert scaler.get_scale() == 65536.0                  # no inf/nan was found
    assert not torch.equal(before, net.head[1].weight)


def test_ingredient_command_captures_default_kwargs():
    """ba3l Ingredient.command adds every default keyword of the decorated factory to the config and later calls the
    factory with config values by NAME (ba3l/ingredients/ingredient.py:84-132, ba3l/module.py:39-40): the keyword names
    and defaults of get_model / AugmentMelSTFT are API.  Stand-in reproducing that capture, fed with the reference's
    own config overrides (ex_audioset.py:61-70)."""
    from passt_b200 import passt as P
    from passt_b200.preprocess import AugmentMelSTFT

    class Ingredient:
        def __init__(self):
            self.config, self.commands = {}, {}

        def command(self, fn):
            sig = inspect.signature(fn.__init__ if inspect.isclass(fn) else fn)
            defaults = {k: v.default for k, v in sig.parameters.items() if v.default is not inspect.Parameter.empty}
            self.config = {**defaults, **self.config}                       # defaults at lowest priority
            def captured(**override):
                kw = {k: self.config[k] for k in defaults}
                kw.update(override)
                return fn(**kw)
            self.commands[fn.__name__] = captured
            return captured

    net_ing = Ingredient()
    net_ing.config.update(arch="passt_s_swa_p16_128_ap476", n_classes=527, s_patchout_t=40, s_patchout_f=4,
                          pretrained=False)
    get_model = net_ing.command(P.get_model)
    assert list(inspect.signature(P.get_model).parameters) == [
        "arch", "pretrained", "n_classes", "in_channels", "fstride", "tstride", "input_fdim", "input_tdim",
        "u_patchout", "s_patchout_t", "s_patchout_f"]                      # models/passt.py:957-961
    mel_ing = Ingredient()
    mel_ing.config.update(n_mels=128, sr=32000, win_length=800, hopsize=320, n_fft=1024, freqm=48, timem=192, htk=False,
                          fmin=0.0, fmax=None, norm=1, fmin_aug_range=10, fmax_aug_range=2000)
    make_mel = mel_ing.command(AugmentMelSTFT)
    with quiet():
        net = get_model().to(DEV).train()
        mel = make_mel().to(DEV).train()
    assert net.s_patchout_t == 40 and net.s_patchout_f == 4 and len(net.blocks) == 12
    net.return_embed = True                                                  # ex_audioset.py:126 sets it from outside
    net = P.lighten_model(net, cut_depth=10)
    x = mel(0.1 * torch.randn(2, 320000, device=DEV))
    y_hat, embed = net(x.unsqueeze(1))                                       # ex_audioset.py:179 unpacking
    assert y_hat.shape == (2, 527) and embed.shape == (2, 768)
