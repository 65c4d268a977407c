"""Waveform-side augmentation on the device: gain, pad/truncate, roll and zero-mean waveform
mixup of the reference's loader workers (audioset/dataset.py:107-140, 315-339) as ONE sm_100a launch that doubles as
the staging pass of a batch (csrc/waveaug.cu).  mp3 decoding stays on the CPU, like in the reference.

    aug = WaveAugment(clip_length=320000, gain_augment=7, roll_range=50, wavmix_rate=0.5, wavmix_beta=2)
    draws = aug.draw(B)                       # host RNG, same distributions / generators as the reference, per clip
    wave, y = aug(raw_list_or_batch, targets, draws)     # [B, L] f32 on the device, mixed targets

CUDA only; there is no CPU path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib as L


@dataclass
class WaveDraws:
    gain: torch.Tensor        # f32 [B] linear amplitude
    shift: torch.Tensor       # int32 [B]
    mix_idx: torch.Tensor     # int32 [B], -1 = not mixed
    mix_lam: torch.Tensor     # f32 [B]


class WaveAugment:
    def __init__(self, clip_length: int = 320000, gain_augment: int = 7, roll_range: int = 50,
                 wavmix_rate: float = 0.0, wavmix_beta: float = 2.0):
        self.clip_length = int(clip_length)
        self.gain_augment = int(gain_augment)
        self.roll_range = int(roll_range)
        self.wavmix_rate = float(wavmix_rate)
        self.wavmix_beta = float(wavmix_beta)

    def draw(self, B: int) -> WaveDraws:
        """Per clip, in the order the reference's wrappers run for one item: gain (torch.randint, dataset.py:112),
        roll (np.random.random_integers(-r, r), :333), wavmix (torch.rand(1) < rate, torch.randint partner,
        np.random.beta, :128-134)."""
        gain = torch.ones(B)
        shift = torch.zeros(B, dtype=torch.int32)
        idx = torch.full((B,), -1, dtype=torch.int32)
        lam = torch.ones(B)
        for b in range(B):
            if self.gain_augment:
                g = torch.randint(self.gain_augment * 2, (1,)).item() - self.gain_augment
                gain[b] = 10 ** (g / 20)
            if self.roll_range:
                shift[b] = int(np.random.randint(-self.roll_range, self.roll_range + 1))
            if self.wavmix_rate > 0 and B > 1 and torch.rand(1).item() < self.wavmix_rate:
                idx[b] = int(torch.randint(B, (1,)).item())
                l = np.random.beta(self.wavmix_beta, self.wavmix_beta)
                lam[b] = max(l, 1.0 - l)
        return WaveDraws(gain, shift, idx, lam)

    def __call__(self, raw: Union[torch.Tensor, Sequence[torch.Tensor]], targets: Optional[torch.Tensor],
                 draws: WaveDraws, out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """raw: device tensor [B, L_in] (fixed-length decode) or a list of B 1-D device tensors (ragged decode)."""
        if isinstance(raw, torch.Tensor):
            if raw.dim() != 2:
                raise ValueError("expected [B, L_in]")
            B, Lin = raw.shape
            flat = raw.detach().float().contiguous()
            offs = torch.arange(B, dtype=torch.int64) * Lin
            lens = torch.full((B,), Lin, dtype=torch.int32)
        else:
            B = len(raw)
            lens = torch.tensor([int(r.numel()) for r in raw], dtype=torch.int32)
            offs = torch.zeros(B, dtype=torch.int64)
            offs[1:] = torch.cumsum(lens[:-1].to(torch.int64), 0)
            flat = torch.cat([r.detach().float().reshape(-1) for r in raw])
        dev = flat.device
        if dev.type != "cuda":
            raise RuntimeError("passt_b200.WaveAugment runs on CUDA (sm_100a) only; there is no CPU path")
        Lc = self.clip_length
        if out is None:
            out = torch.empty(B, Lc, device=dev, dtype=torch.float32)
        mixing = bool((draws.mix_idx >= 0).any())
        small = torch.cat([offs.view(torch.int32).reshape(-1), lens, draws.shift.to(torch.int32),
                           draws.mix_idx.to(torch.int32)]).pin_memory().to(dev, non_blocking=True)
        fl = torch.cat([draws.gain.float(), draws.mix_lam.float()]).pin_memory().to(dev, non_blocking=True)
        d_off, d_len = small[: 2 * B].view(torch.int64), small[2 * B: 3 * B]
        d_shift, d_idx = small[3 * B: 4 * B], small[4 * B: 5 * B]
        d_gain, d_lam = fl[:B], fl[B:]
        tgt = tgt_out = None
        C = 0
        if targets is not None:
            tgt = targets.detach().to(device=dev, dtype=torch.float32).contiguous()
            C = tgt.shape[1]
            tgt_out = torch.empty_like(tgt)
        with torch.cuda.device(dev):
            L.call("passt_wave_augment", L.ptr(flat), L.ptr(d_off), L.ptr(d_len), L.ptr(d_gain), L.ptr(d_shift),
                   L.ptr(d_idx) if mixing else None, L.ptr(d_lam) if mixing else None, L.ptr(out), L.ptr(tgt),
                   L.ptr(tgt_out), B, Lc, C, L.stream_ptr())
        return out, tgt_out
