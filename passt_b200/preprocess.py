"""AugmentMelSTFT — drop-in for the reference frontend module (models/preprocess.py:19-86).

Same constructor keywords and defaults, same ``forward(x[B, L]) -> mel[B, n_mels, T]`` contract, honours
``.training`` (band augmentation + SpecAugment only in train).  The computation is ONE hand-written sm_100a
kernel (passt_b200/csrc/mel.cu) reached through the C ABI; the module only makes the random draws, in the
reference's order and from the reference's generators:
  * ``torch.randint`` x2 on the CPU default generator, in train *and* eval (models/preprocess.py:63-64),
  * ``torch.rand([B])`` x2 per enabled mask on the input's device (torchaudio functional.py:864-865),
and hands them to the kernel as arguments.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib as L


class AugmentMelSTFT(nn.Module):
    def __init__(self, n_mels=128, sr=32000, win_length=800, hopsize=320, n_fft=1024, freqm=48, timem=192,
                 htk=False, fmin=0.0, fmax=None, norm=1, fmin_aug_range=1, fmax_aug_range=1000):
        super().__init__()
        if n_mels != 128 or n_fft != 1024:
            raise ValueError("the sm_100a mel kernel is built for n_mels=128, n_fft=1024 (the PaSST configuration)")
        if win_length > n_fft:
            raise ValueError("win_length must be <= n_fft")
        self.win_length = win_length
        self.n_mels = n_mels
        self.n_fft = n_fft
        self.sr = sr
        self.htk = htk
        self.fmin = fmin
        if fmax is None:
            fmax = sr // 2 - fmax_aug_range // 2          # models/preprocess.py:32-35
            print(f"Warning: FMAX is None setting to {fmax} ")
        self.fmax = fmax
        self.norm = norm
        self.hopsize = hopsize
        # kept for interface compatibility (non-persistent, like the reference :38-46); the kernel builds its own
        # window table on device
        self.register_buffer("window", torch.hann_window(win_length, periodic=False), persistent=False)
        assert fmin_aug_range >= 1, f"fmin_aug_range={fmin_aug_range} should be >=1; 1 means no augmentation"
        assert fmax_aug_range >= 1, f"fmax_aug_range={fmax_aug_range} should be >=1; 1 means no augmentation"
        self.fmin_aug_range = fmin_aug_range
        self.fmax_aug_range = fmax_aug_range
        self.register_buffer("preemphasis_coefficient", torch.as_tensor([[[-.97, 1]]]), persistent=False)
        self.freqm = int(freqm)
        self.timem = int(timem)
        self._ws = {}          # device -> (workspace tensor, last (fmin, fmax))
        self._band_dev = None   # optional float64[2] device buffer: graph replays read (fmin, fmax) from it
        self.last_draws = None  # (fmin, fmax, rnd[4,B] or None) of the most recent call, for parity tests

    def _workspace(self, device):
        key = (device.type, device.index)
        ent = self._ws.get(key)
        if ent is None:
            nbytes = L.load().passt_mel_workspace_bytes()
            ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
            L.call("passt_mel_init", L.ptr(ws), self.win_length, L.stream_ptr())
            ent = [ws, None]
            self._ws[key] = ent
        return ent

    def draw_band(self):
        """The two CPU-generator draws of the reference forward (models/preprocess.py:63-68) -> (fmin, fmax)."""
        r0 = torch.randint(self.fmin_aug_range, (1,)).item()
        r1 = torch.randint(self.fmax_aug_range, (1,)).item()
        if self.training:
            return self.fmin + r0, self.fmax + self.fmax_aug_range // 2 - r1
        return self.fmin, self.fmax

    @torch.compiler.disable
    def forward(self, x, band=None):
        """band: optional (fmin, fmax) drawn by the caller with draw_band() (CUDA-graph replays draw on the host and
        pass the values through a device buffer); None = draw here, like the reference."""
        if not x.is_cuda:
            raise RuntimeError("passt_b200.AugmentMelSTFT runs on CUDA (sm_100a) only; there is no CPU path")
        if x.dim() != 2:
            raise ValueError(f"expected waveform [B, L], got {tuple(x.shape)}")
        x = x.detach()
        if x.dtype != torch.float32:
            x = x.float()
        x = x.contiguous()
        B, Lw = x.shape
        # draw order of the reference: two CPU randints first (also consumed in eval)
        fmin, fmax = band if band is not None else self.draw_band()
        rnd = None
        freqm = timem = 0
        if self.training and (self.freqm > 0 or self.timem > 0):
            parts = []
            zero = None
            if self.freqm > 0:
                parts += [torch.rand(B, device=x.device), torch.rand(B, device=x.device)]
                freqm = self.freqm
            else:
                zero = torch.zeros(B, device=x.device)
                parts += [zero, zero]
            if self.timem > 0:
                parts += [torch.rand(B, device=x.device), torch.rand(B, device=x.device)]
                timem = self.timem
            else:
                zero = zero if zero is not None else torch.zeros(B, device=x.device)
                parts += [zero, zero]
            rnd = torch.stack(parts).contiguous()
        if B == 0:
            # empty batch: same draws as above, empty result (torch.stft-based reference: [0, n_mels, T])
            self.last_draws = (fmin, fmax, rnd)
            return torch.empty(0, self.n_mels, 1 + (Lw - 1) // self.hopsize, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            ent = self._workspace(x.device)
            st = L.stream_ptr()
            if self._band_dev is not None:
                # graph mode: identical launch every step, the band comes from device memory
                L.call("passt_mel_set_band_dev", L.ptr(ent[0]), L.ptr(self._band_dev), int(self.sr), st)
                ent[1] = None
            elif ent[1] != (fmin, fmax):
                L.call("passt_mel_set_band", L.ptr(ent[0]), float(fmin), float(fmax), int(self.sr), st)
                ent[1] = (fmin, fmax)
            T = 1 + (Lw - 1) // self.hopsize
            out = torch.empty(B, self.n_mels, T, device=x.device, dtype=torch.float32)
            L.call("passt_mel_forward", L.ptr(ent[0]), L.ptr(x), L.ptr(out), B, Lw, self.hopsize, L.ptr(rnd), freqm,
                   timem, st)
        self.last_draws = (fmin, fmax, rnd)
        return out

    def extra_repr(self):
        return "winsize={}, hopsize={}".format(self.win_length, self.hopsize)
