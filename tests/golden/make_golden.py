"""Generate tests/golden/passt_golden.pt by running the UNMODIFIED reference (imported from /root/reference through
tests/ref_shim.py) on CPU with reproducible synthetic weights and inputs.  Run in the build container only:

    python tests/golden/make_golden.py

The fixture travels to the GPU box (which has no /root/reference) and pins both the oracle (CPU tests) and the CUDA
path (GPU tests) to real reference outputs.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from ref_shim import load_reference, quiet  # noqa: E402
from oracle import passt_oracle as O  # noqa: E402


def main():
    rp, rpre = load_reference()
    G = {}
    # ---------------- network: structured patchout config (BASELINE config 2 shape, batch 2) -----------------
    net_kw = dict(s_patchout_t=40, s_patchout_f=4)
    cfg = O.NetCfg(**net_kw)
    G.update(net_kw=net_kw, param_seed=1234, input_seed=77, rng_seed=4321, x_shape=(2, 1, 128, 1000))
    with quiet():
        net = rp.get_model(arch="passt_s_swa_p16_128_ap476", pretrained=False, n_classes=527, **net_kw)
    net.load_state_dict(O.synth_params(cfg, seed=G["param_seed"]), strict=True)
    torch.manual_seed(G["input_seed"])
    x = torch.randn(*G["x_shape"])
    net.train()
    torch.manual_seed(G["rng_seed"])
    with quiet():
        logits, feats = net(x)
    torch.manual_seed(G["rng_seed"])
    d = O.draw_patchout(cfg, 12, 99, True)
    G["t_keep"], G["f_keep"] = d.t_keep, d.f_keep
    G["train_logits"] = logits.detach().clone()
    torch.manual_seed(99)
    w = torch.randn_like(logits)
    G["grad_weight_seed"] = 99
    (logits * w).sum().backward()
    grads = {k: p.grad for k, p in net.named_parameters() if p.grad is not None}
    G["grad_absmax"] = {k: float(g.abs().max()) for k, g in grads.items()}
    G["grad_sum"] = {k: float(g.double().sum()) for k, g in grads.items()}
    G["grad_samples"] = {k: g.flatten()[:: max(1, g.numel() // 64)][:64].clone() for k, g in grads.items()}
    net.eval()
    with quiet(), torch.no_grad():
        logits, feats = net(x)
    G["eval_logits"], G["eval_features"] = logits.clone(), feats.clone()
    # ---------------- frontend ----------------------------------------------------------------------------------
    with quiet():
        mel = rpre.AugmentMelSTFT(n_mels=128, sr=32000, win_length=800, hopsize=320, n_fft=1024, freqm=48, timem=192,
                                  htk=False, fmin=0.0, fmax=None, norm=1, fmin_aug_range=10, fmax_aug_range=2000)
    G["wave_seed"], G["wave_shape"] = 5, (2, 48000)
    torch.manual_seed(G["wave_seed"])
    wave = 0.1 * torch.randn(*G["wave_shape"])
    mel.eval()
    with quiet():
        G["mel_eval"] = mel(wave).clone()
    mel.train()
    G["mel_train_seed"] = 31
    torch.manual_seed(G["mel_train_seed"])
    with quiet():
        G["mel_train"] = mel(wave).clone()
    # the draws the reference consumed (CPU generator: 2 randint, then 4 x rand[B])
    torch.manual_seed(G["mel_train_seed"])
    dm = O.draw_mel(O.MelCfg(), True, G["wave_shape"][0])
    G["mel_train_fmin"], G["mel_train_fmax"], G["mel_train_rnd"] = dm.fmin, dm.fmax, dm.mask_rnd
    out = os.path.join(HERE, "passt_golden.pt")
    torch.save(G, out)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
