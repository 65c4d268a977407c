"""Validation / ensemble inference path of the hot path's callers on the sm_100a kernels.

  * ``validation_step(mel, nets, wave, target)`` — ex_audioset.py:216-245: ONE mel for the batch, then every net
    (``net`` and, with SWA, ``net_swa``) on that same spectrogram, per-net BCE loss (fused kernel) and
    ``sigmoid(logits)`` (fused kernel); nothing goes to the host.
  * ``EnsembleRunner`` — EnsembelerModel (models/passt.py:1021-1036): logit mean of several nets (different strides
    see different patch grids of the SAME mel) and its sigmoid in one launch.
  * ``MeanAPMeter`` — validation_epoch_end (ex_audioset.py:247-266): accumulates ``out`` / ``target`` on the device
    and computes per-class average precision there (sklearn.metrics.average_precision_score semantics incl. ties),
    replacing the ``.cpu()`` + sklearn hop.
CUDA only; there is no CPU path.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib as L
from . import loss as fused_loss


def _sigmoid_mean(logit_list: Sequence[torch.Tensor], mode: int = 0, want_mean_logits: bool = False):
    z0 = logit_list[0]
    if not z0.is_cuda:
        raise RuntimeError("passt_b200.evalpath runs on CUDA (sm_100a) only; there is no CPU path")
    zs = [z.detach().float().contiguous() for z in logit_list]
    for z in zs:
        if z.shape != zs[0].shape:
            raise ValueError("all nets must produce logits of the same shape")
    K = len(zs)
    if K > 16:
        raise ValueError("at most 16 nets per ensemble launch")
    prob = torch.empty_like(zs[0])
    mean = torch.empty_like(zs[0]) if want_mean_logits else None
    arr = (ctypes.c_void_p * K)(*[z.data_ptr() for z in zs])
    with torch.cuda.device(z0.device):
        L.call("passt_ens_sigmoid", ctypes.cast(arr, ctypes.c_void_p), K, L.ptr(mean), L.ptr(prob), zs[0].numel(), mode,
               L.stream_ptr())
    return (prob, mean) if want_mean_logits else prob


@torch.no_grad()
def validation_step(mel, nets: Sequence[Tuple[str, torch.nn.Module]], wave: torch.Tensor,
                    target: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """nets: [("", net), ("swa_", net_swa)] like ex_audioset.py:231-233.  wave [B, L] (or [B, 1, L]) on the device.
    Returns {prefix+"val_loss", prefix+"out", prefix+"target"} per net, all device tensors."""
    if wave.dim() == 3:
        wave = wave.reshape(-1, wave.shape[-1])                   # mel_forward's reshape (ex_audioset.py:142-146)
    spec = mel(wave).unsqueeze(1)                                 # shared by every net
    res: Dict[str, torch.Tensor] = {}
    for prefix, net in nets:
        logits, _ = net(spec)
        if target is not None:
            res[prefix + "val_loss"] = fused_loss.bce_with_logits(logits, target)
            res[prefix + "target"] = target
        res[prefix + "out"] = _sigmoid_mean([logits])
    return res


class EnsembleRunner(torch.nn.Module):
    """Logit average of several nets on one spectrogram (models/passt.py:1021-1036) + optional sigmoid, one fused
    launch after the nets.  ``forward(x)`` returns ``(mean_logits, mean_logits)`` like EnsembelerModel."""

    def __init__(self, models: Sequence[torch.nn.Module]):
        super().__init__()
        self.models = torch.nn.ModuleList(models)

    def forward(self, x):
        outs = [m(x)[0] for m in self.models]
        _, mean = _sigmoid_mean(outs, mode=0, want_mean_logits=True)
        return mean, mean

    @torch.no_grad()
    def predict_proba(self, x):
        return _sigmoid_mean([m(x)[0] for m in self.models], mode=0)


class MeanAPMeter:
    """Device-side accumulation of validation outputs and per-class AP at epoch end."""

    def __init__(self):
        self.out: List[torch.Tensor] = []
        self.target: List[torch.Tensor] = []

    def update(self, out: torch.Tensor, target: torch.Tensor):
        self.out.append(out.detach().float())
        self.target.append(target.detach().float())

    @torch.no_grad()
    def average_precision(self) -> torch.Tensor:
        out = torch.cat(self.out, 0).contiguous()
        tgt = torch.cat(self.target, 0).contiguous()
        if not out.is_cuda:
            raise RuntimeError("passt_b200.evalpath runs on CUDA (sm_100a) only; there is no CPU path")
        n, C = out.shape
        ap = torch.empty(C, device=out.device, dtype=torch.float32)
        with torch.cuda.device(out.device):
            L.call("passt_average_precision", L.ptr(out), L.ptr(tgt), L.ptr(ap), n, C, L.stream_ptr())
        return ap

    def mean_ap(self) -> torch.Tensor:
        return self.average_precision().mean()      # NaN if a class has no positive, like the reference's nan array

    def reset(self):
        self.out, self.target = [], []
