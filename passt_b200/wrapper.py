"""hear21passt-style convenience surface named by the reference README (README.md:49-85):

    model = get_basic_model(mode="logits")      # .mel (frontend) + .net (PaSST), callable on wave[B, samples]
    model.net = get_model_passt(arch="passt_s_swa_p16_128_ap476", n_classes=50)

The hear21passt package itself is not part of the reference tree; only the calls shown in the README are pinned.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .passt import get_model
from .preprocess import AugmentMelSTFT


class PasstBasicWrapper(nn.Module):
    """wave [B, L] @32 kHz -> logits (mode="logits"), embeddings (mode="embed_only") or both concatenated
    (mode="all"), through the fused mel kernel and the PaSST engine."""

    def __init__(self, mel: nn.Module, net: nn.Module, mode: str = "logits"):
        super().__init__()
        if mode not in ("logits", "embed_only", "all"):
            raise RuntimeError(f"mode='{mode}' is not recognized not in: all, embed_only, logits")
        self.mel = mel
        self.net = net
        self.mode = mode

    def forward(self, x):
        if x.dim() == 3 and x.shape[1] == 1:
            x = x[:, 0]
        specs = self.mel(x)                      # [B, 128, T]
        specs = specs.unsqueeze(1)               # [B, 1, 128, T]  (ex_audioset.py:142-153 mel_forward)
        logits, features = self.net(specs)
        if self.mode == "logits":
            return logits
        if self.mode == "embed_only":
            return features
        return torch.cat([logits, features], dim=1)


def get_model_passt(arch="passt_s_swa_p16_128_ap476", pretrained=True, n_classes=527, in_channels=1, fstride=10,
                    tstride=10, input_fdim=128, input_tdim=998, u_patchout=0, s_patchout_t=0, s_patchout_f=0):
    return get_model(arch=arch, pretrained=pretrained, n_classes=n_classes, in_channels=in_channels, fstride=fstride,
                     tstride=tstride, input_fdim=input_fdim, input_tdim=input_tdim, u_patchout=u_patchout,
                     s_patchout_t=s_patchout_t, s_patchout_f=s_patchout_f)


def get_basic_model(mode="logits", arch="passt_s_swa_p16_128_ap476", pretrained=True, **kwargs):
    mel = AugmentMelSTFT(n_mels=128, sr=32000, win_length=800, hopsize=320, n_fft=1024, freqm=48, timem=192,
                         htk=False, fmin=0.0, fmax=None, norm=1, fmin_aug_range=10, fmax_aug_range=2000)
    net = get_model_passt(arch=arch, pretrained=pretrained, **kwargs)
    return PasstBasicWrapper(mel=mel, net=net, mode=mode)
