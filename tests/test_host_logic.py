"""CPU tests of the host side: drop-in surface, RNG draw order, state_dict contract, C-ABI exports, error behaviour."""
import ctypes
import os
import re

import pytest
import torch

from util import ROOT, quiet
from oracle import passt_oracle as O


def test_state_dict_keys_and_shapes_match_reference_contract():
    from passt_b200.passt import get_model
    for arch, depth in [("passt_s_swa_p16_128_ap476", 12), ("passt_l_kd_p16_128_ap47", 7)]:
        with quiet():
            net = get_model(arch=arch, pretrained=False, n_classes=527)
        want = O.param_shapes(O.NetCfg(depth=depth))
        got = {k: tuple(v.shape) for k, v in net.state_dict().items()}
        assert got == want
    assert sum(p.numel() for p in net.parameters()) == 50714398     # passt_l


def test_param_count_passt_s():
    from passt_b200.passt import get_model
    with quiet():
        net = get_model(arch="passt_s_kd_p16_128_ap486", pretrained=False)
    assert sum(p.numel() for p in net.parameters()) == 86153758
    assert net.no_weight_decay() == {"new_pos_embed", "freq_new_pos_embed", "time_new_pos_embed", "cls_token",
                                     "dist_token"}
    assert net.patch_embed.grid_size == (12, 99) and net.num_tokens == 2


def test_get_model_errors_and_strides():
    from passt_b200.passt import get_model
    with pytest.raises(RuntimeError, match="Unknown model"):
        get_model(arch="nope", pretrained=False)
    with pytest.raises(RuntimeError, match="pretrained=True"):
        get_model(arch="passt_s_swa_p16_128_ap476")           # no network / no checkpoint dir
    with quiet():
        net = get_model(arch="passt_s_p16_s16_128_ap468", pretrained=False, fstride=16, tstride=16)
    assert net.patch_embed.grid_size == (8, 62)
    with quiet():
        net = get_model(arch="passt_s_f128_30sec_p16_s10_ap473", pretrained=False, input_tdim=3000)
    assert net.patch_embed.grid_size == (12, 300)


@pytest.mark.parametrize("kw,T,training", [
    (dict(s_patchout_t=40, s_patchout_f=4), 1000, True), (dict(u_patchout=400), 1000, True),
    (dict(s_patchout_t=10, s_patchout_f=3), 500, True), (dict(), 1000, False), (dict(), 500, False),
    (dict(s_patchout_t=40, s_patchout_f=4, u_patchout=100), 998, True)])
def test_step_plan_draws_match_oracle(kw, T, training):
    """Host RNG order + token bookkeeping: bit-exact patchout indices / offset vs the oracle's restatement."""
    from passt_b200 import engine
    from passt_b200.passt import get_model
    with quiet():
        net = get_model(arch="passt_s_swa_p16_128_ap476", pretrained=False, **kw)
    cfg = O.NetCfg(**kw)
    x = torch.zeros(2, 1, 128, T)
    torch.manual_seed(123)
    plan = engine.draw_step_plan(net, x, training)
    after_a = torch.rand(1)
    torch.manual_seed(123)
    fg, tg = O.conv_grid(cfg, 128, T)
    d = O.draw_patchout(cfg, fg, tg, training)
    after_b = torch.rand(1)
    assert torch.equal(after_a, after_b)                 # same number of generator draws consumed
    assert plan.toffset == d.toffset
    for a, b in ((plan.t_keep, d.t_keep), (plan.f_keep, d.f_keep), (plan.u_keep, d.u_keep)):
        assert (a is None) == (b is None) and (a is None or torch.equal(a, b))
    assert plan.ntok == O.token_count(cfg, 128, T, training)
    # token order: F-major, T-minor, then unstructured selection — same gather as the oracle on an index image
    tg_eff = min(tg, cfg.grid[1])
    idx = torch.arange(fg * tg_eff).reshape(1, 1, fg, tg_eff).float()
    z = idx
    if d.t_keep is not None:
        z = z[:, :, :, d.t_keep]
    if d.f_keep is not None:
        z = z[:, :, d.f_keep, :]
    z = z.flatten(2)[0, 0]
    if d.u_keep is not None:
        z = z[d.u_keep]
    mine = plan.patch_f.long() * tg_eff + plan.patch_t.long()
    assert torch.equal(mine, z.long())


def test_mel_module_draw_order_without_gpu():
    from passt_b200.preprocess import AugmentMelSTFT
    with quiet():
        mel = AugmentMelSTFT(fmin_aug_range=10, fmax_aug_range=2000)
    assert mel.fmax == 15000 and "window" not in mel.state_dict()
    with pytest.raises(RuntimeError, match="CUDA"):
        mel(torch.zeros(1, 32000))                       # no CPU fallback, fails loudly


def test_no_cpu_fallback_in_net():
    from passt_b200.passt import get_model
    with quiet():
        net = get_model(arch="passt_s_swa_p16_128_ap476", pretrained=False)
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.zeros(1, 1, 128, 1000))
    with pytest.raises(RuntimeError, match="parameter container"):
        net.blocks[0](torch.zeros(1, 4, 768))


def test_dropin_models_namespace():
    import subprocess, sys
    code = ("import sys; sys.path.insert(0, %r); import models.passt as mp, models.preprocess as mq; "
            "assert hasattr(mp, 'model_ing') and hasattr(mq, 'model_ing'); "
            "assert callable(mp.get_model) and callable(mp.get_ensemble_model) and callable(mp.lighten_model); "
            "print('ok')") % os.path.join(ROOT, "passt_b200", "dropin")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT,
                         env={**os.environ, "PYTHONPATH": ROOT})
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr


def test_lighten_and_deepcopy():
    import copy
    from passt_b200.passt import get_model, lighten_model
    with quiet():
        net = get_model(arch="passt_s_swa_p16_128_ap476", pretrained=False)
    net2 = copy.deepcopy(net)                            # SWA deep-copies the net (swa_callback.py:140)
    assert net2._wcache is not net._wcache
    assert all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), net2.state_dict().values()))
    lighten_model(net2, cut_depth=9)
    assert len(net2.blocks) == 3


def test_c_abi_library_exports_every_declared_symbol():
    from passt_b200 import _lib, build
    lib_path = build.build()
    lib = ctypes.CDLL(lib_path)
    header = open(os.path.join(ROOT, "include", "passt_b200.h")).read()
    declared = set(re.findall(r"\b(passt_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 18
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/passt_b200.h but not exported"
    for sym in _lib.exported_symbols():
        assert sym in declared, f"{sym} bound in _lib.py but not declared in the header"


def test_fused_adamw_refuses_cpu_parameters():
    """No CPU fallback anywhere on the product path: the optimizer says so instead of silently running elsewhere."""
    import pytest
    import torch
    from passt_b200.optim import FusedAdamW
    with pytest.raises(RuntimeError, match="CUDA"):
        FusedAdamW([torch.nn.Parameter(torch.zeros(8))], lr=1e-3)
