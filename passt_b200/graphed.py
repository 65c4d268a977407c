"""CUDA-graph replay of one full training step (frontend -> PaSST forward -> loss -> hand-written backward ->
optional gradient all-reduce -> optimizer), for launch-bound steps: ~300 kernel launches become one graph launch.

Everything that changes from step to step is data, not launch arguments:
  * the waveform batch and targets are copied into static device buffers,
  * the host-side random draws the reference makes (mel band, time-embedding offset, patchout indices — same torch
    CPU-generator calls in the same order, so indices stay bit-identical to the reference for a seed) are written
    into small static device buffers that the kernels read (passt_mel_set_band_dev, token_table toff_dev),
  * SpecAugment uniforms come from torch.rand inside the capture (graph-safe Philox offsets).

    step = GraphedTrainStep(mel, net, optimizer, loss_fn, example_wave, example_target)
    loss = step(wave, target)            # wave/target: device or pinned-host tensors of the example's shape

The optimizer must be graph-safe: ``passt_b200.optim.FusedAdamW`` or a torch optimizer created with
``capturable=True``.  Parity with the eager step: tests/test_gpu_graphed.py.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import engine


class GraphedTrainStep:
    def __init__(self, mel, net, optimizer, loss_fn: Callable, example_wave: torch.Tensor,
                 example_target: torch.Tensor, reducer=None, warmup: int = 3, restore_after_warmup: bool = True,
                 mixup_alpha: Optional[float] = None):
        """Building the step runs ``max(1, warmup)`` REAL eager training steps on the example batch (allocator,
        cudaFuncSetAttribute, bf16 weight cache, optimizer state and pointer tables must exist before capture).
        With ``restore_after_warmup`` (default) the network parameters and the optimizer state are put back to what
        they were before those steps (fresh optimizer state is zeroed), so constructing the object has no training
        side effect; pass False to keep the warm-up updates.
        ``mixup_alpha``: spectrogram mixup of the reference's training_step (ex_audioset.py:172-177) -- the permutation
        and the lambdas are drawn per step on the host (helpers/mixup.py order: after the mel draws, before the
        network's), written into static device buffers, folded into the patch gather (PaSST.fused_mixup) and passed
        to ``loss_fn(logits, target, perm, lam)`` (passt_b200.loss.bce_with_logits / cross_entropy mix the targets)."""
        if not example_wave.is_cuda:
            raise RuntimeError("GraphedTrainStep needs CUDA tensors (sm_100a); there is no CPU path")
        self.mel, self.net, self.opt, self.loss_fn, self.reducer = mel, net, optimizer, loss_fn, reducer
        dev = example_wave.device
        self.wave = torch.empty_like(example_wave)
        self.target = torch.empty_like(example_target)
        self.band_dev = torch.zeros(2, dtype=torch.float64, device=dev)
        self.toff_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        mel.train(); net.train()
        B, Lw = example_wave.shape
        self.mixup_alpha = mixup_alpha
        self.perm_dev = torch.zeros(B, dtype=torch.int32, device=dev) if mixup_alpha else None
        self.lam_dev = torch.ones(B, dtype=torch.float32, device=dev) if mixup_alpha else None
        T = 1 + (Lw - 1) // mel.hopsize
        self._x_shape = (B, 1, mel.n_mels, T)
        # static index buffer sized from one trial plan
        probe = engine.draw_step_plan(net, torch.empty(self._x_shape, device=dev), True)
        self.idx_dev = torch.zeros(2, probe.ntok - 2, dtype=torch.int32, device=dev)
        self.graph = None
        self.loss = None
        snap = self._snapshot() if restore_after_warmup else None
        # eager warm-up on a side stream (allocator, cudaFuncSetAttribute, bf16 weight cache, optimizer state)
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self.wave.copy_(example_wave); self.target.copy_(example_target)
            for _ in range(max(1, warmup)):
                self._host_draws()
                self._body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self._host_draws()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._body()
        torch.cuda.synchronize()
        if snap is not None:
            self._restore(snap)

    # ---- state handling around the warm-up / after a checkpoint load ------------------------------------------
    def _snapshot(self):
        params = [p.detach().clone() for p in self.net.parameters()]
        had_state = len(self.opt.state) > 0
        opt_state = None
        if had_state:
            opt_state = [{k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in self.opt.state[p].items()}
                         if p in self.opt.state else None
                         for g in self.opt.param_groups for p in g["params"]]
        return params, had_state, opt_state

    @torch.no_grad()
    def _restore(self, snap):
        params, had_state, opt_state = snap
        for p, v in zip(self.net.parameters(), params):
            p.copy_(v)                               # in place: the captured graph keeps the parameter addresses
        flat = [p for g in self.opt.param_groups for p in g["params"]]
        for i, p in enumerate(flat):
            st = self.opt.state.get(p)
            if not st:
                continue
            for k, v in st.items():
                if not torch.is_tensor(v):
                    continue
                old = opt_state[i].get(k) if (had_state and opt_state[i] is not None) else None
                if old is not None:
                    v.copy_(old)
                else:
                    v.zero_()                         # state created by the warm-up: back to "never stepped"
        self.resync()

    def resync(self):
        """Call after the parameters were changed behind the graph's back (``net.load_state_dict``, manual edits):
        re-casts the bf16 GEMM-operand copies the captured forward reads.  (With FusedAdamW attached the captured
        forward contains no refresh of its own -- the optimizer pass rewrites the copies -- so without this the first
        replay after a checkpoint load would run on stale bf16 weights.)"""
        depth = len(self.net.blocks)
        P = dict(self.net.named_parameters())
        wnames = ["patch_embed.proj.weight"] + [f"blocks.{i}.{w}.weight" for i in range(depth)
                                                for w in ("attn.qkv", "attn.proj", "mlp.fc1", "mlp.fc2")]
        with torch.cuda.device(self.wave.device):
            self.net._wcache.refresh_all([P[n] for n in wnames])

    # host RNG, in the reference's order: mel band first (preprocess.py:63-64), then the network's draws
    def _host_draws(self):
        fmin, fmax = self.mel.draw_band()
        # a FRESH pinned temporary per call (like draw_step_plan's index copies): the caching host allocator keeps the
        # block alive until the queued copy has run, so a host that runs ahead of the GPU can never overwrite the band
        # of a step that has not executed yet (a persistent pinned mirror would be torn / reused across steps)
        self.band_dev.copy_(torch.tensor([float(fmin), float(fmax)], dtype=torch.float64).pin_memory(),
                            non_blocking=True)
        if self.mixup_alpha:
            from .loss import draw_mixup
            perm, lam = draw_mixup(self._x_shape[0], self.mixup_alpha)
            self.perm_dev.copy_(perm.to(torch.int32).pin_memory(), non_blocking=True)
            self.lam_dev.copy_(lam.pin_memory(), non_blocking=True)
        x_meta = torch.empty(self._x_shape, device="meta")
        plan = engine.draw_step_plan(self.net, _ShapeOnCuda(x_meta, self.wave.device), True, static_idx=self.idx_dev,
                                     static_toff=self.toff_dev)
        self._plan = plan
        self._band = (fmin, fmax)

    def _body(self):
        self.mel._band_dev = self.band_dev
        self.net._preset_plan = self._plan
        try:
            with torch.no_grad():
                spec = self.mel(self.wave, band=self._band).unsqueeze(1)
            if self.mixup_alpha:
                self.net.fused_mixup(self.perm_dev, self.lam_dev)
                logits, _ = self.net(spec)
                loss = self.loss_fn(logits, self.target, self.perm_dev, self.lam_dev)
            else:
                logits, _ = self.net(spec)
                loss = self.loss_fn(logits, self.target)
            self.opt.zero_grad(set_to_none=True)
            loss.backward()
            if self.reducer is not None:
                self.reducer.all_reduce()
            self.opt.step()
        finally:
            self.mel._band_dev = None
            self.net._preset_plan = None
        return loss

    @staticmethod
    def _stage(dst: torch.Tensor, src: torch.Tensor):
        """Copy a step input into its static buffer.  Device-resident sources are copied by a kernel, not by
        cudaMemcpy: a device-to-device memcpy can be queued on the copy engine behind a loader's host-to-device
        transfer of the *next* batch and then stalls the whole step for the length of that transfer (measured:
        +1.1 ms per step with an 82 MB prefetch in flight)."""
        if src.is_cuda and src.dtype == dst.dtype and src.shape == dst.shape:
            torch.mul(src, 1, out=dst)
        else:
            dst.copy_(src, non_blocking=True)

    def __call__(self, wave: Optional[torch.Tensor] = None, target: Optional[torch.Tensor] = None,
                 consumed: Optional[torch.cuda.Event] = None):
        """consumed: optional event recorded as soon as the inputs have been copied out of ``wave`` / ``target`` --
        a prefetching loader may refill those buffers from then on, it does not have to wait for the step."""
        if wave is not None:
            self._stage(self.wave, wave)
        if target is not None:
            self._stage(self.target, target)
        if consumed is not None:
            consumed.record(torch.cuda.current_stream())
        self._host_draws()
        if hasattr(self.opt, "sync_hyperparams"):
            self.opt.sync_hyperparams()       # passt_b200.FusedAdamW: the replayed step reads lr etc. from a pinned mirror
        self.graph.replay()
        self.net._wcache.dirty = True     # the replay updated the parameters: the next eager forward must re-cast
        return self.loss


class GraphedInference:
    """CUDA-graph replay of the inference path (waveform -> mel (eval) -> PaSST forward): ~110 launches become one.
    ``logits = step(wave)``; the returned tensor is the graph's static output buffer (valid until the next call).
    The fp32 parameters other than the GEMM weights (LayerNorm, biases, positional embeddings, head) are read in place;
    after changing weights call ``resync()``."""

    def __init__(self, mel, net, example_wave: torch.Tensor, warmup: int = 2):
        if not example_wave.is_cuda:
            raise RuntimeError("GraphedInference needs CUDA tensors (sm_100a); there is no CPU path")
        self.mel, self.net = mel, net
        mel.eval(); net.eval()
        self.wave = torch.empty_like(example_wave)
        dev = example_wave.device
        B, Lw = example_wave.shape
        x_shape = (B, 1, mel.n_mels, 1 + (Lw - 1) // mel.hopsize)
        # the (deterministic) eval token plan lives in static device buffers: a plan drawn inside the capture would leave
        # a memcpy node that re-reads a pinned temporary long after it was freed
        probe = engine.draw_step_plan(net, _ShapeOnCuda(torch.empty(x_shape, device="meta"), dev), False)
        self.idx_dev = torch.zeros(2, probe.ntok - 2, dtype=torch.int32, device=dev)
        self.toff_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self._plan = engine.draw_step_plan(net, _ShapeOnCuda(torch.empty(x_shape, device="meta"), dev), False,
                                           static_idx=self.idx_dev, static_toff=self.toff_dev)
        torch.cuda.synchronize()
        s = torch.cuda.Stream(device=example_wave.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self.wave.copy_(example_wave)
            for _ in range(max(1, warmup)):
                self._body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.logits, self.features = self._body()
        torch.cuda.synchronize()

    def resync(self):
        """Call after the parameters were changed behind the graph's back (``load_state_dict``, SWA averaging into this
        net, manual edits): re-casts the bf16 GEMM-operand copies the captured forward reads (the capture holds no
        refresh of its own: the warm-up passes had already cleared the cache's dirty flag)."""
        depth = len(self.net.blocks)
        P = dict(self.net.named_parameters())
        wnames = ["patch_embed.proj.weight"] + [f"blocks.{i}.{w}.weight" for i in range(depth)
                                                for w in ("attn.qkv", "attn.proj", "mlp.fc1", "mlp.fc2")]
        with torch.cuda.device(self.wave.device):
            self.net._wcache.refresh_all([P[n] for n in wnames])

    def _body(self):
        with torch.no_grad():
            # eval: the band is fixed and no patchout / offset draws reach the kernels, but the two CPU randint draws of
            # the reference's mel forward (preprocess.py:63-64) are still consumed -- per replay, in __call__
            spec = self.mel(self.wave, band=(self.mel.fmin, self.mel.fmax)).unsqueeze(1)
            self.net._preset_plan = self._plan
            try:
                return self.net(spec)
            finally:
                self.net._preset_plan = None

    def __call__(self, wave: Optional[torch.Tensor] = None, consumed: Optional[torch.cuda.Event] = None):
        if wave is not None:
            GraphedTrainStep._stage(self.wave, wave)
        if consumed is not None:
            consumed.record(torch.cuda.current_stream())
        self.mel.draw_band()                 # same CPU-generator consumption as an eager eval forward
        self.graph.replay()
        return self.logits


class _ShapeOnCuda:
    """draw_step_plan only needs the input's shape and device."""

    def __init__(self, meta, device):
        self.shape = meta.shape
        self.device = device
        self.is_cuda = True
