"""FusedAdamW — the optimizer step of the training hot path as ONE hand-written sm_100a launch.

Drop-in for the optimizer the reference builds (ex_audioset.py:104-109: ``torch.optim.AdamW(params, lr=lr,
weight_decay=weight_decay)``): same constructor keywords and defaults, same update rule, ``state_dict`` with the usual
``step / exp_avg / exp_avg_sq`` entries.  What differs is where the work happens:

  * all parameters of a group are updated by a single kernel (``passt_adamw_step``) driven by a device-resident
    pointer table; the two moments live in two flat fp32 buffers (views per parameter),
  * the step count lives on the device and the learning rate is read from device memory, so a CUDA-graph replay of
    the whole train step (``passt_b200.graphed.GraphedTrainStep``) advances both correctly
    (call ``sync_hyperparams()`` after an LR scheduler changed ``group["lr"]`` -- GraphedTrainStep does),
  * ``attach(net)``: weights that have a bf16 GEMM-operand copy in the network's weight cache get that copy rewritten
    in the same pass, and the network is told to skip its own per-step refresh.

CUDA only; there is no CPU path.
"""
from __future__ import annotations

import torch

from . import _lib as L


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        if lr < 0.0 or eps < 0.0 or weight_decay < 0.0 or not (0.0 <= betas[0] < 1.0 and 0.0 <= betas[1] < 1.0):
            raise ValueError("invalid AdamW hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._net = None
        self._g = []          # per param group: pointer tables, pinned/device hyper-parameters, flat moment buffers
        for group in self.param_groups:
            ps = [p for p in group["params"]]
            if not ps:
                self._g.append(None)
                continue
            dev = ps[0].device
            if dev.type != "cuda":
                raise RuntimeError("passt_b200.FusedAdamW runs on CUDA (sm_100a) only; there is no CPU path")
            offs, total = [], 0
            for p in ps:
                if p.dtype != torch.float32 or not p.is_contiguous() or p.device != dev:
                    raise ValueError("FusedAdamW expects contiguous fp32 parameters on one CUDA device")
                offs.append(total)
                total += (p.numel() + 3) // 4 * 4          # keep every moment view 16-byte aligned
            m = torch.zeros(total, device=dev)
            v = torch.zeros(total, device=dev)
            hyper_dev = torch.zeros(8, device=dev)
            for p, o in zip(ps, offs):
                self.state[p] = dict(step=hyper_dev[5], exp_avg=m[o:o + p.numel()].view_as(p),
                                     exp_avg_sq=v[o:o + p.numel()].view_as(p))
            self._g.append(dict(tables={}, hyper_dev=hyper_dev, m=m, v=v, offs=offs))
        self.sync_hyperparams()

    # ---- hyper-parameters live in device memory (hyper_dev[0:5]; [5] is the step count).  They are pushed with a
    #      stream-ordered copy from a FRESH pinned temporary per call (the caching host allocator keeps it alive until
    #      the copy has executed), never from a persistent pinned mirror: a host that runs ahead of the GPU could
    #      overwrite such a mirror with the lr of step N+k before step N's copy has run.  Eager step() pushes by itself;
    #      a captured step contains no copy at all -- GraphedTrainStep calls sync_hyperparams() right before each replay.
    def sync_hyperparams(self):
        for group, g in zip(self.param_groups, self._g):
            if g is None:
                continue
            h = torch.tensor([float(group["lr"]), float(group["betas"][0]), float(group["betas"][1]),
                              float(group["eps"]), float(group["weight_decay"])], dtype=torch.float32)
            dev = g["hyper_dev"].device
            with torch.cuda.device(dev):
                g["hyper_dev"][:5].copy_(h.pin_memory(), non_blocking=True)

    @staticmethod
    def _drop_uncaptured(tables):
        """Forget pointer tables, EXCEPT those a CUDA graph was captured with: the graph's memcpy node re-reads that
        pinned host table on every replay, so freeing it would let the pinned allocator hand the block to someone else
        (e.g. a DataLoader's pin_memory) and the kernel would dereference garbage."""
        for k in [k for k, e in tables.items() if not e[4]]:
            tables.pop(k)

    def attach(self, net):
        """Let the step also rewrite the bf16 copies of ``net``'s weight matrices (passt_b200.PaSST)."""
        self._net = net
        for g in self._g:
            if g is not None:
                self._drop_uncaptured(g["tables"])
        return self

    def _bf16_copy(self, p):
        if self._net is None:
            return None
        ent = self._net._wcache._store.get((p.data_ptr(), tuple(p.shape)))
        if ent is None or ent[2] is not None:      # no copy yet, or a transposed copy is kept too: leave to the cache
            return None
        return ent[1]

    def _table(self, group, g):
        """Device pointer table for the parameters that currently have gradients.  One (pinned host, device) pair is
        kept per distinct pointer set and never rewritten: a captured CUDA graph keeps reading the pair of its capture
        while eager steps in between (whose gradient buffers live elsewhere) use their own."""
        ps = [p for p in group["params"] if p.grad is not None]
        recs = []
        for p in ps:
            if p.grad.dtype != torch.float32 or not p.grad.is_contiguous():
                raise RuntimeError("FusedAdamW expects contiguous fp32 gradients")
            st = self.state[p]
            wb = self._bf16_copy(p)
            recs.append((p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                         0 if wb is None else wb.data_ptr(), p.numel()))
        sig = tuple(recs)
        ent = g["tables"].get(sig)
        if ent is None:
            if not recs:
                return None
            rows, blk = [], 0
            for pp, gp, mp, vp, wp, n in recs:
                vec = int(n % 4 == 0 and all(a % 16 == 0 for a in (pp, gp, mp, vp)) and wp % 8 == 0)
                rows.append([pp, gp, mp, vp, wp, n, blk | (vec << 32), 0])
                blk += (n + 4095) // 4096
            host = torch.tensor(rows, dtype=torch.int64).reshape(-1, 8).pin_memory()
            dev = torch.empty(host.shape, dtype=torch.int64, device=g["hyper_dev"].device)
            # pinned -> device: legal inside a capture (a memcpy node that re-reads this never-modified pinned table)
            dev.copy_(host, non_blocking=True)
            # bound the cache when eager gradient buffers keep moving; a table a CUDA graph was captured with is never
            # dropped (the graph's memcpy node and kernel keep reading it)
            evictable = [k for k, e in g["tables"].items() if not e[4]]
            if len(evictable) >= 8:
                g["tables"].pop(evictable[0])
            ent = (host, dev, blk, len(recs), torch.cuda.is_current_stream_capturing())
            g["tables"][sig] = ent
        return ent

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        refreshed = False
        if not torch.cuda.is_current_stream_capturing():
            self.sync_hyperparams()
        for group, g in zip(self.param_groups, self._g):
            if g is None:
                continue
            ent = self._table(group, g)
            if ent is None:
                continue
            _, table_dev, blocks, n, _ = ent
            dev = g["hyper_dev"].device
            with torch.cuda.device(dev):
                L.call("passt_adamw_step", L.ptr(table_dev), n, blocks, L.ptr(g["hyper_dev"]), L.stream_ptr())
            refreshed = True
        if refreshed and self._net is not None:
            self._net._wcache.fresh_from_optimizer = True
        return loss

    def load_state_dict(self, state_dict):
        """Values are copied into the flat moment buffers (the kernel's pointer tables keep pointing at them)."""
        super().load_state_dict(state_dict)
        for group, g in zip(self.param_groups, self._g):
            if g is None:
                continue
            self._drop_uncaptured(g["tables"])
            for p, o in zip(group["params"], g["offs"]):
                st = self.state.get(p)
                if not st:
                    continue
                mv = g["m"][o:o + p.numel()].view_as(p)
                vv = g["v"][o:o + p.numel()].view_as(p)
                mv.copy_(st["exp_avg"]); vv.copy_(st["exp_avg_sq"])
                g["hyper_dev"][5] = float(st["step"])
                self.state[p] = dict(step=g["hyper_dev"][5], exp_avg=mv, exp_avg_sq=vv)
        self.sync_hyperparams()
