"""Data-parallel gradient exchange for the PaSST engine (one process per GPU, NCCL over NVLink/NVSwitch).

The reference trains with pytorch-lightning DDP (ex_audioset.py:475-496): a bucketed NCCL all-reduce of all
parameter gradients each step, with ``head_dist.*`` never receiving a gradient (models/passt.py:582-588).
Here the hand-written backward writes every gradient into ONE flat fp32 buffer in parameter order
(engine.PasstFunction.backward); as soon as the backward pass has finished a contiguous chunk of it (classifier
tail, then each transformer block from last to first, then the patch-embedding head) the chunk's all-reduce is
enqueued asynchronously, so the exchange overlaps the remaining backward kernels.  The path shards by batch and
this is its only collective.

Contract with autograd (checked, not assumed): the backward returns VIEWS of the flat buffer as the parameter
gradients, and averaging is only correct if ``p.grad`` ends up aliasing that buffer -- i.e. the gradients did not
exist before the backward (``optimizer.zero_grad(set_to_none=True)``, the torch default).  With a pre-existing
``.grad`` (``set_to_none=False``, gradient accumulation) AccumulateGrad would add the still un-reduced view into a
different tensor while NCCL is reducing the flat buffer; ``all_reduce()`` detects that and raises instead of letting
the replicas diverge silently.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


def suggest_nccl_ctas(n_param_bytes: int, world: int, backward_seconds: float, per_cta_gbs: float = 16.0,
                      overlap_fraction: float = 0.6) -> int:
    """How many SMs (= NCCL CTAs, ``NCCL_MAX_CTAS``) the gradient all-reduce needs so that it finishes inside the part
    of the backward pass it can overlap with.  A ring all-reduce moves ``2 (world-1)/world`` x the gradient bytes over
    each GPU's links; one NCCL CTA sustained ~16 GB/s next to the persistent tcgen05 kernels here (measured on 2 x B200:
    345 MB in ~5.5 ms with 4 CTAs).  Too few CTAs expose the all-reduce at small per-GPU batches (cfg5: 67 % weak-scaling
    efficiency with 4), too many take SMs from every backward GEMM (cfg2: 95.0 % with 8 vs 97.0 % with 4).
    Clamped to [4, 32].  With 4 or more ranks NCCL reduces inside the NVSwitch (NVLS, 4 channels) and is fastest left
    alone: at 8 x B200 the step took 23.63 ms unreserved vs 24.00 / 24.06 / 24.33 ms with 4 / 8 / 16 reserved SMs
    (profiles/r2_bench_scale8_*.json), so 0 (= no reservation, NCCL's own CTA count) is returned there."""
    if world <= 1 or world >= 4:
        return 0
    wire = 2.0 * (world - 1) / world * n_param_bytes
    need = wire / (per_cta_gbs * 1e9 * overlap_fraction * max(backward_seconds, 1e-4))
    return int(min(32, max(4, -(-need // 1))))


class GradAllReducer:
    def __init__(self, net: Optional[torch.nn.Module] = None, group=None, average: bool = True,
                 min_chunk_elems: int = 4 * 1024 * 1024, reserve_sms: int = 4):
        """reserve_sms: SMs the backward's persistent kernels (GEMMs, attention) leave free while the chunk all-reduces
        are in flight (engine.PasstFunction.backward lowers passt_set_sm_limit by this much).  A 148-CTA persistent
        GEMM cannot share an SM with an NCCL CTA (shared memory), so without the reservation every GEMM that overlaps
        an all-reduce waits for whole SMs; run NCCL with NCCL_MAX_CTAS <= reserve_sms (bench.py sets it)."""
        self.reserve_sms = int(reserve_sms)
        self.group = group
        self.average = average
        self.min_chunk = min_chunk_elems
        self._works: List = []
        self._pending = None      # (flat, lo, hi) chunk being coalesced
        self._net = net
        self._flat = None         # flat gradient buffer of the backward pass being reduced
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if net is not None:
            net._grad_chunk_hook = self.on_chunk_ready

    # called by the backward pass: flat[lo:hi] holds final gradient values
    def on_chunk_ready(self, flat: torch.Tensor, lo: int, hi: int):
        if self.world == 1 or hi <= lo:
            return
        self._flat = flat
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
        if self._net is not None and self._flat is not None:
            self.check_grads_alias(self._net, self._flat)
        self._flat = None

    @staticmethod
    def check_grads_alias(net, flat=None):
        """Raise unless every parameter gradient of ``net`` lives inside the flat buffer that was all-reduced."""
        flat = flat if flat is not None else getattr(net, "_last_flat_grad", None)
        if flat is None:
            return
        lo = flat.data_ptr()
        hi = lo + flat.numel() * flat.element_size()
        for n, p in net.named_parameters():
            g = p.grad
            if g is None:
                continue
            if not (lo <= g.data_ptr() < hi):
                raise RuntimeError(
                    f"passt_b200.ddp: the gradient of '{n}' does not alias the all-reduced flat buffer, so it holds "
                    "un-reduced values (a .grad existed before backward: use optimizer.zero_grad(set_to_none=True), "
                    "and do not accumulate gradients over several backward passes with GradAllReducer)")

    finish = all_reduce
