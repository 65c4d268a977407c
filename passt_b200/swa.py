"""Stochastic weight averaging of the PaSST net as one sm_100a launch.

Replaces ``StochasticWeightAveraging.update_parameters`` / ``avg_fn`` (helpers/swa_callback.py:246-268): a Python loop
over ~160 parameter pairs doing ``p_swa.copy_(p_swa + (p - p_swa)/(n+1))`` (3 ATen kernels each) becomes a single
multi-tensor kernel driven by a device pointer table.  The averaged copy is a ``deepcopy`` of the net, exactly like
the reference (swa_callback.py:140), so it can be evaluated next to the live net (passt_b200.evalpath).  The
reference averages at epoch ends (not per step), so this is deliberately NOT folded into the per-step AdamW pass.
"""
from __future__ import annotations

import copy

import torch

from . import _lib as L


class SWAAverager:
    def __init__(self, net: torch.nn.Module, net_swa: torch.nn.Module = None):
        self.net = net
        self.net_swa = net_swa if net_swa is not None else copy.deepcopy(net)     # swa_callback.py:140
        for p in self.net_swa.parameters():
            p.requires_grad_(False)
        self.n_averaged = 0
        self._table = None
        self._sig = None

    def _build_table(self):
        pairs = list(zip(self.net_swa.parameters(), self.net.parameters()))
        sig = tuple((a.data_ptr(), b.data_ptr(), a.numel()) for a, b in pairs)
        if sig == self._sig:
            return
        rows, blk = [], 0
        dev = pairs[0][0].device
        if dev.type != "cuda":
            raise RuntimeError("passt_b200.SWAAverager runs on CUDA (sm_100a) only; there is no CPU path")
        for a, b in pairs:
            if a.shape != b.shape or a.dtype != torch.float32 or b.dtype != torch.float32 or a.device != b.device:
                raise ValueError("SWAAverager expects two fp32 parameter lists of identical shapes on one device")
            if not (a.is_contiguous() and b.is_contiguous()):
                raise ValueError("SWAAverager expects contiguous parameters")
            rows.append([b.data_ptr(), a.data_ptr(), a.numel(), blk])
            blk += (a.numel() + 4095) // 4096
        self._table = torch.tensor(rows, dtype=torch.int64).to(dev)
        self._blocks, self._n, self._sig = blk, len(rows), sig

    @torch.no_grad()
    def update(self):
        """One averaging step (call where the reference's callback does: at the end of an epoch >= swa_epoch_start)."""
        self._build_table()
        dev = self._table.device
        with torch.cuda.device(dev):
            L.call("passt_swa_update", L.ptr(self._table), self._n, self._blocks, int(self.n_averaged), L.stream_ptr())
        self.n_averaged += 1
        wc = getattr(self.net_swa, "_wcache", None)
        if wc is not None:
            wc.invalidate()          # the kernel does not bump Tensor._version: force the bf16 copies to be re-cast
        return self.net_swa
