"""Data-parallel gradient exchange for the PaSST engine (one process per GPU, NCCL over NVLink/NVSwitch).

The reference trains with pytorch-lightning DDP (ex_audioset.py:475-496): a bucketed NCCL all-reduce of all
parameter gradients each step, with ``head_dist.*`` never receiving a gradient (models/passt.py:582-588).
Here the hand-written backward writes every gradient into ONE flat fp32 buffer in parameter order
(engine.PasstFunction.backward); as soon as the backward pass has finished a contiguous chunk of it (classifier
tail, then each transformer block from last to first, then the patch-embedding head) the chunk's all-reduce is
enqueued asynchronously, so the exchange overlaps the remaining backward kernels.  The path shards by batch and
this is its only collective.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


class GradAllReducer:
    def __init__(self, net: Optional[torch.nn.Module] = None, group=None, average: bool = True,
                 min_chunk_elems: int = 4 * 1024 * 1024):
        self.group = group
        self.average = average
        self.min_chunk = min_chunk_elems
        self._works: List = []
        self._pending = None      # (flat, lo, hi) chunk being coalesced
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if net is not None:
            net._grad_chunk_hook = self.on_chunk_ready

    # called by the backward pass: flat[lo:hi] holds final gradient values
    def on_chunk_ready(self, flat: torch.Tensor, lo: int, hi: int):
        if self.world == 1 or hi <= lo:
            return
        if self._pending is not None:
            pf, plo, phi = self._pending
            if pf is flat and (hi == plo or lo == phi):
                lo, hi = min(lo, plo), max(hi, phi)      # coalesce with the adjacent pending chunk
            else:
                self._launch(pf, plo, phi)
            self._pending = None
        if hi - lo < self.min_chunk:
            self._pending = (flat, lo, hi)
        else:
            self._launch(flat, lo, hi)

    def _launch(self, flat, lo, hi):
        chunk = flat[lo:hi]
        if self.average and dist.get_backend(self.group) == "nccl":
            w = dist.all_reduce(chunk, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
            self._works.append((w, None))
        else:
            w = dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._works.append((w, chunk if self.average else None))

    def all_reduce(self):
        """Flush pending chunks and make the current stream wait for every outstanding all-reduce."""
        if self._pending is not None:
            self._launch(*self._pending)
            self._pending = None
        for w, chunk in self._works:
            w.wait()
            if chunk is not None:
                chunk.div_(self.world)
        self._works = []

    finish = all_reduce
