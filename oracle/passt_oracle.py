"""CPU oracle for the PaSST hot path  —  TEST INFRASTRUCTURE ONLY.

This file is a *functional restatement* (plain torch CPU ops, fp32) of the two reference modules on the hot
path.  It exists to check the CUDA kernels; nothing in the product package (``passt_b200/``) may import it.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs use it.

Pinning status: the reference repository ships no tests, golden vectors or fixtures for this path
("parity unpinned" by the reference's own tests).  The oracle is therefore pinned against
the *reference implementation itself*, imported from /root/reference through import shims
(``tests/ref_shim.py``): ``tests/test_oracle_vs_reference.py`` compares every function here with the reference
modules on seeded inputs (bit-exact for indices, <=1e-6 for floats), and ``tests/golden/make_golden.py``
commits reference outputs as fixtures that travel to the GPU box.

Where the arithmetic lives in third-party code (torch / torchaudio, pinned by the reference to
torch 1.11-1.13 / torchaudio 0.11-0.13; this image has 2.11), the published algorithm is restated here and the
call sites are cited.

Reference citations are relative to /root/reference (kkoutini/PaSST @ 2a5c818).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------------
# configuration records
# --------------------------------------------------------------------------------------------------
@dataclass
class MelCfg:
    """Constructor arguments of AugmentMelSTFT (models/preprocess.py:20-21) as wired by ex_audioset.py:64-69."""
    n_mels: int = 128
    sr: int = 32000
    win_length: int = 800
    hopsize: int = 320
    n_fft: int = 1024
    freqm: int = 48
    timem: int = 192
    fmin: float = 0.0
    fmax: Optional[float] = None
    fmin_aug_range: int = 10
    fmax_aug_range: int = 2000

    def resolved_fmax(self) -> float:
        # models/preprocess.py:32-35
        return self.fmax if self.fmax is not None else self.sr // 2 - self.fmax_aug_range // 2


@dataclass
class NetCfg:
    """Arguments of get_model / PaSST.__init__ (models/passt.py:957-961, :391-396)."""
    depth: int = 12
    embed_dim: int = 768
    num_heads: int = 12
    mlp_ratio: float = 4.0
    patch: int = 16
    fstride: int = 10
    tstride: int = 10
    input_fdim: int = 128
    input_tdim: int = 998
    n_classes: int = 527
    u_patchout: int = 0
    s_patchout_t: int = 0
    s_patchout_f: int = 0

    @property
    def grid(self) -> Tuple[int, int]:
        # PatchEmbed.grid_size = img_size // stride (models/passt.py:311)
        return self.input_fdim // self.fstride, self.input_tdim // self.tstride


@dataclass
class StepDraws:
    """Every random quantity one forward pass consumes, in the reference's draw order."""
    fmin: float = 0.0
    fmax: float = 15000.0
    mask_rnd: Optional[torch.Tensor] = None   # [4, B] uniforms: freq value, freq min, time value, time min
    toffset: int = 0
    t_keep: Optional[torch.Tensor] = None     # sorted kept time columns (int64) or None
    f_keep: Optional[torch.Tensor] = None     # sorted kept freq rows or None
    u_keep: Optional[torch.Tensor] = None     # sorted kept flattened positions or None
    extra: dict = field(default_factory=dict)


# --------------------------------------------------------------------------------------------------
# frontend
# --------------------------------------------------------------------------------------------------
def kaldi_mel_banks(n_mels: int, n_fft: int, sr: float, fmin: float, fmax: float) -> torch.Tensor:
    """Triangular mel filterbank [n_mels, n_fft//2], vtln_warp == 1.0.

    Restates torchaudio.compliance.kaldi.get_mel_banks (site-packages/torchaudio/compliance/kaldi.py:436-511)
    as called at models/preprocess.py:71-72.  mel(f) = 1127 ln(1 + f/700); endpoints are python floats.
    """
    nyquist = 0.5 * sr
    if fmax <= 0.0:
        fmax += nyquist
    n_bins = n_fft // 2
    bin_hz = sr / n_fft
    m_lo = 1127.0 * math.log(1.0 + fmin / 700.0)
    m_hi = 1127.0 * math.log(1.0 + fmax / 700.0)
    step = (m_hi - m_lo) / (n_mels + 1)
    idx = torch.arange(n_mels).unsqueeze(1)
    left = m_lo + idx * step
    center = m_lo + (idx + 1.0) * step
    right = m_lo + (idx + 2.0) * step
    mel_of_bin = (1127.0 * (1.0 + (bin_hz * torch.arange(float(n_bins))) / 700.0).log()).unsqueeze(0)
    rising = (mel_of_bin - left) / (center - left)
    falling = (right - mel_of_bin) / (right - center)
    return torch.clamp_min(torch.minimum(rising, falling), 0.0)


def draw_mel(cfg: MelCfg, training: bool, batch: int, device="cpu") -> StepDraws:
    """RNG consumption of AugmentMelSTFT.forward: two CPU randint draws happen in train AND eval
    (models/preprocess.py:63-64); eval then overrides with the fixed band (:66-68).  The four SpecAugment
    uniforms come from the generator of the spectrogram's device (torchaudio functional.py:864-865)."""
    d = StepDraws()
    r0 = torch.randint(cfg.fmin_aug_range, (1,)).item()
    r1 = torch.randint(cfg.fmax_aug_range, (1,)).item()
    if training:
        d.fmin = cfg.fmin + r0
        d.fmax = cfg.resolved_fmax() + cfg.fmax_aug_range // 2 - r1
        rnd = []
        if cfg.freqm > 0:
            rnd += [torch.rand(batch, device=device), torch.rand(batch, device=device)]
        else:
            rnd += [torch.zeros(batch, device=device)] * 2
        if cfg.timem > 0:
            rnd += [torch.rand(batch, device=device), torch.rand(batch, device=device)]
        else:
            rnd += [torch.zeros(batch, device=device)] * 2
        d.mask_rnd = torch.stack(rnd)
    else:
        d.fmin = cfg.fmin
        d.fmax = cfg.resolved_fmax()
    return d


def mel_power_spectrum(wave: torch.Tensor, cfg: MelCfg) -> torch.Tensor:
    """[B, L] -> power spectrogram [B, n_fft/2+1, T].

    Restates models/preprocess.py:59-62: pre-emphasis y[n] = x[n+1] - 0.97 x[n]; torch.stft with center=True
    (reflect pad n_fft/2), hann(win_length, periodic=False) zero-padded to n_fft and centred; |.|^2.
    """
    y = wave[:, 1:] - 0.97 * wave[:, :-1]
    half = cfg.n_fft // 2
    y = F.pad(y.unsqueeze(1), (half, half), mode="reflect").squeeze(1)
    frames = y.unfold(-1, cfg.n_fft, cfg.hopsize)                      # [B, T, n_fft]
    win = torch.hann_window(cfg.win_length, periodic=False, dtype=wave.dtype, device=wave.device)
    lpad = (cfg.n_fft - cfg.win_length) // 2
    win = F.pad(win, (lpad, cfg.n_fft - cfg.win_length - lpad))
    spec = torch.fft.rfft(frames * win, dim=-1)                        # [B, T, n_fft/2+1]
    return (spec.real ** 2 + spec.imag ** 2).transpose(1, 2)


def band_mask(x: torch.Tensor, u_value: torch.Tensor, u_min: torch.Tensor, mask_param: int, axis: int):
    """torchaudio.functional.mask_along_axis_iid (functional.py:813-882) with mask_value 0.0, p = 1.0."""
    size = x.shape[axis]
    value = u_value * mask_param
    mn = u_min * (size - value)
    start = mn.long()
    end = mn.long() + value.long()
    pos = torch.arange(size, device=x.device)
    shape = [1, 1, 1]
    shape[axis] = size
    pos = pos.view(shape)
    hit = (pos >= start.view(-1, 1, 1)) & (pos < end.view(-1, 1, 1))
    return x.masked_fill(hit, 0.0)


def mel_frontend(wave: torch.Tensor, cfg: MelCfg, draws: StepDraws, training: bool) -> torch.Tensor:
    """AugmentMelSTFT.forward (models/preprocess.py:57-86): [B, L] -> [B, n_mels, T]."""
    power = mel_power_spectrum(wave, cfg)
    bank = kaldi_mel_banks(cfg.n_mels, cfg.n_fft, cfg.sr, draws.fmin, draws.fmax)
    bank = F.pad(bank, (0, 1)).to(wave.device)                          # zero Nyquist column (:73)
    mel = torch.matmul(bank, power)                                     # (:76)
    mel = (mel + 0.00001).log()                                         # (:78)
    if training and draws.mask_rnd is not None:
        if cfg.freqm > 0:
            mel = band_mask(mel, draws.mask_rnd[0], draws.mask_rnd[1], cfg.freqm, 1)   # (:81)
        if cfg.timem > 0:
            mel = band_mask(mel, draws.mask_rnd[2], draws.mask_rnd[3], cfg.timem, 2)   # (:82)
    return (mel + 4.5) / 5.0                                            # (:84)


# --------------------------------------------------------------------------------------------------
# network
# --------------------------------------------------------------------------------------------------
def param_shapes(cfg: NetCfg) -> Dict[str, Tuple[int, ...]]:
    """state_dict keys and shapes of PaSST (models/passt.py:428-467), distilled=True."""
    Dm, H = cfg.embed_dim, int(cfg.embed_dim * cfg.mlp_ratio)
    Fg, Tg = cfg.grid
    s: Dict[str, Tuple[int, ...]] = {
        "cls_token": (1, 1, Dm), "dist_token": (1, 1, Dm), "new_pos_embed": (1, 2, Dm),
        "freq_new_pos_embed": (1, Dm, Fg, 1), "time_new_pos_embed": (1, Dm, 1, Tg),
        "patch_embed.proj.weight": (Dm, 1, cfg.patch, cfg.patch), "patch_embed.proj.bias": (Dm,),
    }
    for i in range(cfg.depth):
        p = f"blocks.{i}."
        s[p + "norm1.weight"] = (Dm,); s[p + "norm1.bias"] = (Dm,)
        s[p + "attn.qkv.weight"] = (3 * Dm, Dm); s[p + "attn.qkv.bias"] = (3 * Dm,)
        s[p + "attn.proj.weight"] = (Dm, Dm); s[p + "attn.proj.bias"] = (Dm,)
        s[p + "norm2.weight"] = (Dm,); s[p + "norm2.bias"] = (Dm,)
        s[p + "mlp.fc1.weight"] = (H, Dm); s[p + "mlp.fc1.bias"] = (H,)
        s[p + "mlp.fc2.weight"] = (Dm, H); s[p + "mlp.fc2.bias"] = (Dm,)
    s["norm.weight"] = (Dm,); s["norm.bias"] = (Dm,)
    s["head.0.weight"] = (Dm,); s["head.0.bias"] = (Dm,)
    s["head.1.weight"] = (cfg.n_classes, Dm); s["head.1.bias"] = (cfg.n_classes,)
    s["head_dist.weight"] = (cfg.n_classes, Dm); s["head_dist.bias"] = (cfg.n_classes,)
    return s


def synth_params(cfg: NetCfg, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Deterministic, reference-independent parity weights: every tensor is perturbed (no structural zeros), so
    each gradient path is exercised (SURVEY.md §8c caveat).  Reproducible on any box with the same torch."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in param_shapes(cfg).items():
        if name.endswith("norm1.weight") or name.endswith("norm2.weight") or name in ("norm.weight", "head.0.weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            t = 0.02 * torch.randn(shape, generator=g)
        elif name == "patch_embed.proj.weight":
            t = 0.05 * torch.randn(shape, generator=g)
        elif "pos_embed" in name or "token" in name:
            t = 0.02 * torch.randn(shape, generator=g)
        else:
            t = 0.02 * torch.randn(shape, generator=g)
        out[name] = t
    return out


def draw_patchout(cfg: NetCfg, f_dim: int, t_dim: int, training: bool) -> StepDraws:
    """CPU-generator draws of PaSST.forward_features, in order (models/passt.py:516, :535, :541, :551)."""
    d = StepDraws()
    t_embed = cfg.grid[1]
    t_eff = t_dim
    if t_dim < t_embed:
        if training:
            d.toffset = torch.randint(1 + t_embed - t_dim, (1,)).item()
    else:
        t_eff = t_embed                                       # x is cut to the embedding length (:523-526)
    if training and cfg.s_patchout_t:
        d.t_keep = torch.randperm(t_eff)[: t_eff - cfg.s_patchout_t].sort().values
    if training and cfg.s_patchout_f:
        d.f_keep = torch.randperm(f_dim)[: f_dim - cfg.s_patchout_f].sort().values
    if training and cfg.u_patchout:
        nt = len(d.t_keep) if d.t_keep is not None else t_eff
        nf = len(d.f_keep) if d.f_keep is not None else f_dim
        seq = nt * nf
        d.u_keep = torch.randperm(seq)[: seq - cfg.u_patchout].sort().values
    return d


def conv_grid(cfg: NetCfg, f_in: int, t_in: int) -> Tuple[int, int]:
    """Conv2d output size (models/passt.py:315): (dim - patch)//stride + 1."""
    return (f_in - cfg.patch) // cfg.fstride + 1, (t_in - cfg.patch) // cfg.tstride + 1


def tokens_from_mel(p: Dict[str, torch.Tensor], x: torch.Tensor, cfg: NetCfg, d: StepDraws) -> torch.Tensor:
    """forward_features part 1 (models/passt.py:506-564): [B,1,F,T] -> [B,N,D] token sequence."""
    B = x.shape[0]
    z = F.conv2d(x, p["patch_embed.proj.weight"], p["patch_embed.proj.bias"], stride=(cfg.fstride, cfg.tstride))
    t_pos = p["time_new_pos_embed"]
    if z.shape[-1] < t_pos.shape[-1]:
        t_pos = t_pos[:, :, :, d.toffset: d.toffset + z.shape[-1]]
    else:
        z = z[:, :, :, : t_pos.shape[-1]]
    z = z + t_pos
    z = z + p["freq_new_pos_embed"]
    if d.t_keep is not None:
        z = z[:, :, :, d.t_keep]
    if d.f_keep is not None:
        z = z[:, :, d.f_keep, :]
    z = z.flatten(2).transpose(1, 2)                           # F-major, T-minor token order (:546)
    if d.u_keep is not None:
        z = z[:, d.u_keep, :]
    cls = p["cls_token"].expand(B, -1, -1) + p["new_pos_embed"][:, :1, :]
    dist = p["dist_token"].expand(B, -1, -1) + p["new_pos_embed"][:, 1:, :]
    return torch.cat((cls, dist, z), dim=1)


def block_forward(p: Dict[str, torch.Tensor], i: int, x: torch.Tensor, cfg: NetCfg) -> torch.Tensor:
    """Block.forward (models/passt.py:377-380) with Attention (:343-360) and Mlp (:283-289); dropouts are p=0."""
    pre = f"blocks.{i}."
    B, N, C = x.shape
    H = cfg.num_heads
    h = F.layer_norm(x, (C,), p[pre + "norm1.weight"], p[pre + "norm1.bias"], 1e-6)
    qkv = F.linear(h, p[pre + "attn.qkv.weight"], p[pre + "attn.qkv.bias"])
    qkv = qkv.reshape(B, N, 3, H, C // H).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    att = (q @ k.transpose(-2, -1)) * ((C // H) ** -0.5)
    att = att.softmax(dim=-1)
    a = (att @ v).transpose(1, 2).reshape(B, N, C)
    x = x + F.linear(a, p[pre + "attn.proj.weight"], p[pre + "attn.proj.bias"])
    h = F.layer_norm(x, (C,), p[pre + "norm2.weight"], p[pre + "norm2.bias"], 1e-6)
    h = F.gelu(F.linear(h, p[pre + "mlp.fc1.weight"], p[pre + "mlp.fc1.bias"]))
    return x + F.linear(h, p[pre + "mlp.fc2.weight"], p[pre + "mlp.fc2.bias"])


def passt_forward(p: Dict[str, torch.Tensor], x: torch.Tensor, cfg: NetCfg, d: StepDraws):
    """PaSST.forward (models/passt.py:576-595): [B,1,F,T] -> (logits [B,C], features [B,D])."""
    tok = tokens_from_mel(p, x, cfg, d)
    for i in range(cfg.depth):
        tok = block_forward(p, i, tok, cfg)
    C = tok.shape[-1]
    tok = F.layer_norm(tok, (C,), p["norm.weight"], p["norm.bias"], 1e-6)          # (:570)
    features = (tok[:, 0] + tok[:, 1]) / 2                                          # (:583)
    hl = F.layer_norm(features, (C,), p["head.0.weight"], p["head.0.bias"], 1e-5)   # head = LN + Linear (:463-464)
    logits = F.linear(hl, p["head.1.weight"], p["head.1.bias"])
    return logits, features


def token_count(cfg: NetCfg, f_in: int, t_in: int, training: bool) -> int:
    fg, tg = conv_grid(cfg, f_in, t_in)
    tg = min(tg, cfg.grid[1])
    if training:
        return (fg - cfg.s_patchout_f) * (tg - cfg.s_patchout_t) - cfg.u_patchout + 2
    return fg * tg + 2
