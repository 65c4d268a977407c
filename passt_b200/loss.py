"""Fused training losses of the hot path's callers, each ONE sm_100a launch forward and ONE
backward (csrc/loss.cu), instead of the ~8-kernel ATen chain of ``F.binary_cross_entropy_with_logits`` plus the
target-mixing elementwise ops:

  * ``bce_with_logits(logits, target, perm=None, lam=None)``  — training_step of ex_audioset.py:172-192
    (``y_mix = y*lam + y[perm]*(1-lam)``; mean BCE over B x C);
  * ``cross_entropy(logits, target, perm=None, lam=None)``    — ex_esc50.py:151-169
    (mean of ``CE(z, y)*lam + CE(z, y[perm])*(1-lam)``).

The kernel writes the loss scalar and d loss / d logits together; ``backward`` multiplies by the upstream gradient,
which it reads from device memory (no host sync, CUDA-graph safe).  ``draw_mixup`` makes the reference's random
draws (helpers/mixup.py:5-12) in the reference's order.  CUDA only; there is no CPU path.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib as L

_WS = {}


def _workspace(dev: torch.device, B: int) -> torch.Tensor:
    key = (dev.index, B)
    ws = _WS.get(key)
    if ws is None:
        ws = torch.zeros(L.load().passt_loss_workspace_bytes(B), dtype=torch.uint8, device=dev)
        _WS[key] = ws
    return ws


def draw_mixup(size: int, alpha: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """helpers/mixup.py:5-12, same generators in the same order: ``torch.randperm`` (torch CPU generator), then
    ``np.random.beta`` (numpy global RNG); lam = max(l, 1-l).  Returns CPU tensors (perm int64 [size], lam f32 [size])."""
    perm = torch.randperm(size)
    lambd = np.random.beta(alpha, alpha, size).astype(np.float32)
    lambd = np.concatenate([lambd[:, None], 1 - lambd[:, None]], 1).max(1)
    return perm, torch.from_numpy(lambd)


class _FusedLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, perm, lam, kind):
        if not logits.is_cuda:
            raise RuntimeError("passt_b200.loss runs on CUDA (sm_100a) only; there is no CPU path")
        if logits.dim() != 2:
            raise ValueError(f"expected logits [B, C], got {tuple(logits.shape)}")
        z = logits.detach()
        if z.dtype != torch.float32:
            z = z.float()
        z = z.contiguous()
        B, C = z.shape
        dev = z.device
        if B == 0:
            raise ValueError("empty batch")
        if kind == "bce":
            t = target.detach().to(device=dev, dtype=torch.float32).contiguous()
            if t.shape != z.shape:
                raise ValueError(f"target shape {tuple(t.shape)} != logits shape {tuple(z.shape)}")
        else:
            t = target.detach().to(device=dev, dtype=torch.int64).contiguous()
            if t.shape != (B,):
                raise ValueError(f"expected class indices [B], got {tuple(t.shape)}")
        if (perm is None) != (lam is None):
            raise ValueError("perm and lam come together")
        if perm is not None:
            perm = perm.detach().to(device=dev, dtype=torch.int32).contiguous()
            lam = lam.detach().to(device=dev, dtype=torch.float32).contiguous()
        loss = torch.empty(1, device=dev, dtype=torch.float32)
        need_grad = ctx.needs_input_grad[0]
        dl = torch.empty_like(z) if need_grad else None
        with torch.cuda.device(dev):
            L.call("passt_loss_bce" if kind == "bce" else "passt_loss_ce", L.ptr(z), L.ptr(t), L.ptr(perm), L.ptr(lam),
                   L.ptr(loss), L.ptr(dl), L.ptr(_workspace(dev, B)), B, C, L.stream_ptr())
        ctx.dl = dl
        ctx.in_dtype = logits.dtype
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gout):
        dl = ctx.dl
        g = gout.detach().to(dtype=torch.float32).reshape(1).contiguous()
        out = torch.empty_like(dl)
        with torch.cuda.device(dl.device):
            L.call("passt_scale_dev", L.ptr(out), L.ptr(dl), L.ptr(g), dl.numel(), L.stream_ptr())
        if ctx.in_dtype != torch.float32:
            out = out.to(ctx.in_dtype)
        return out, None, None, None, None


def bce_with_logits(logits: torch.Tensor, target: torch.Tensor, perm: Optional[torch.Tensor] = None,
                    lam: Optional[torch.Tensor] = None) -> torch.Tensor:
    """mean(BCEWithLogits(logits, target*lam + target[perm]*(1-lam))) — identical to the unmixed loss when perm is None."""
    return _FusedLoss.apply(logits, target, perm, lam, "bce")


def cross_entropy(logits: torch.Tensor, target: torch.Tensor, perm: Optional[torch.Tensor] = None,
                  lam: Optional[torch.Tensor] = None) -> torch.Tensor:
    """mean(CE(logits, target)*lam + CE(logits, target[perm])*(1-lam)); plain mean CE when perm is None."""
    return _FusedLoss.apply(logits, target, perm, lam, "ce")
