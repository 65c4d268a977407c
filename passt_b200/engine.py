"""Host-side orchestration of the sm_100a kernels for one PaSST forward/backward.

This is plumbing only: it owns no arithmetic.  Every tensor op on the path is a call into libpasst_b200.so
(include/passt_b200.h) on torch's current CUDA stream; torch supplies device memory and autograd glue.

Reference being replaced: PaSST.forward_features / forward (models/passt.py:506-595), Block / Attention / Mlp
(:271-380) and their autograd.  Random draws (time-embedding offset :516, structured patchout :535/:541,
unstructured patchout :551) are made here with the *same torch CPU-generator calls in the same order* so that
patchout indices are bit-identical to the reference for a given seed.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from . import _lib as L

BF16 = torch.bfloat16


@dataclass
class StepPlan:
    """Everything one forward pass needs that is decided on the host."""
    B: int
    ntok: int                 # tokens per clip including cls + dist
    Fm: int                   # mel bins of the input
    Tm: int                   # mel frames of the input
    toffset: int
    patch_f: torch.Tensor     # int32 [ntok-2] device: freq grid row of each kept patch
    patch_t: torch.Tensor     # int32 [ntok-2] device: time grid column of each kept patch
    toffset_dev: Optional[torch.Tensor] = None   # int32[1] device copy of toffset (graph replays read it on device)
    t_keep: Optional[torch.Tensor] = None   # CPU int64 draws (exposed for parity tests)
    f_keep: Optional[torch.Tensor] = None
    u_keep: Optional[torch.Tensor] = None
    grad_mode: bool = True    # torch.is_grad_enabled() at the call site (Function.forward always runs with it off, and
                              # ctx.needs_input_grad ignores it): no activations are kept under torch.no_grad()


def draw_step_plan(net, x: torch.Tensor, training: bool, static_idx: Optional[torch.Tensor] = None,
                   static_toff: Optional[torch.Tensor] = None) -> StepPlan:
    """Host RNG + index bookkeeping of forward_features (models/passt.py:508-553).
    static_idx (int32 [2, ntok-2]) / static_toff (int32 [1]): device buffers to refill in place (CUDA-graph replays
    must see the new draws at the same addresses); fresh buffers are allocated when they are None."""
    B, Cin, Fm, Tm = x.shape
    ps, (fs, ts) = net.patch_size, net.stride
    f_dim = (Fm - ps) // fs + 1                       # Conv2d output size (:315)
    t_dim = (Tm - ps) // ts + 1
    Fg, Tg = net.patch_embed.grid_size
    if f_dim != Fg:
        raise RuntimeError(f"input has {f_dim} patch rows but the frequency embedding has {Fg} (passt.py:529)")
    toffset = 0
    if t_dim < Tg:
        if training:
            toffset = int(torch.randint(1 + Tg - t_dim, (1,)).item())       # (:516)
    else:
        t_dim = Tg                                                          # x is cut to the embedding (:523-526)
    t_idx = torch.arange(t_dim)
    f_idx = torch.arange(f_dim)
    t_keep = f_keep = u_keep = None
    if training and net.s_patchout_t:
        t_keep = torch.randperm(t_dim)[: t_dim - net.s_patchout_t].sort().values   # (:535)
        t_idx = t_keep
    if training and net.s_patchout_f:
        f_keep = torch.randperm(f_dim)[: f_dim - net.s_patchout_f].sort().values   # (:541)
        f_idx = f_keep
    # flatten(2): F-major, T-minor token order (:546)
    pf = f_idx.repeat_interleave(len(t_idx))
    pt = t_idx.repeat(len(f_idx))
    if training and net.u_patchout:
        seq = pf.numel()
        u_keep = torch.randperm(seq)[: seq - net.u_patchout].sort().values         # (:551)
        pf, pt = pf[u_keep], pt[u_keep]
    host = torch.stack([pf, pt]).to(torch.int32)
    if x.is_cuda:
        host = host.pin_memory()
    toff_dev = None
    if static_idx is not None:
        static_idx.copy_(host, non_blocking=True)
        dev = static_idx
        if static_toff is not None:
            static_toff.copy_(torch.tensor([toffset], dtype=torch.int32).pin_memory(), non_blocking=True)
            toff_dev = static_toff
    else:
        dev = host.to(x.device, non_blocking=True)
    return StepPlan(B=B, ntok=pf.numel() + 2, Fm=Fm, Tm=Tm, toffset=toffset, patch_f=dev[0], patch_t=dev[1],
                    toffset_dev=toff_dev, t_keep=t_keep, f_keep=f_keep, u_keep=u_keep)


class WeightCache:
    """bf16 copies of the fp32 master weights W [out,in] (forward: K-major B operand; dgrad: the same bytes read as a
    [K, N] row-major MN-major B operand; a transposed copy is only produced on request), refreshed by one cast
    kernel per matrix.

    Staleness cannot be detected from ``Tensor._version``: fused CUDA optimizers (torch.optim.AdamW(fused=True))
    and CUDA-graph replays update parameters without bumping it.  Policy: every *training* forward refreshes
    unconditionally (one multi-matrix cast launch, ``refresh_all``) and marks the cache dirty; the next no-grad forward
    refreshes once more (the optimizer ran after the last training forward) and clears the flag.  In-place edits
    that do bump the version (load_state_dict, SWA averaging, manual ``p.data`` changes via ops) are caught by the
    version check as well."""

    def __init__(self):
        self._store: Dict[tuple, tuple] = {}
        self.gen = 0              # bumped whenever the parameters may have changed (dirty <- True)
        self._dirty = False
        # set by passt_b200.optim.FusedAdamW after a step that rewrote the bf16 copies itself: the next forward skips
        # its own refresh once
        self.fresh_from_optimizer = False

    @property
    def dirty(self) -> bool:
        return self._dirty

    @dirty.setter
    def dirty(self, v: bool):
        self._dirty = bool(v)
        if v:
            self.gen += 1

    def get(self, p: torch.Tensor, need_t: bool, force: bool = False):
        key = (p.data_ptr(), tuple(p.shape))
        ent = self._store.get(key)
        ver = (p.data_ptr(), p._version)
        if (not force and ent is not None and ent[0] == ver and (ent[2] is not None or not need_t)):
            return ent[1], ent[2]
        w2 = p.detach().reshape(p.shape[0], -1)
        R, C = w2.shape
        wb = ent[1] if ent is not None and ent[1].device == p.device else torch.empty(R, C, dtype=BF16, device=p.device)
        wt = ent[2] if ent is not None and ent[2] is not None and ent[2].device == p.device else None
        if need_t and wt is None:
            wt = torch.empty(C, R, dtype=BF16, device=p.device)
        # an existing transposed copy is always kept and refreshed too: a captured CUDA graph may hold its address
        L.call("passt_cast_transpose", L.ptr(w2), L.ptr(wb), L.ptr(wt), R, C, L.stream_ptr())
        self._store[key] = (ver, wb, wt)
        return wb, wt

    def refresh_all(self, params):
        """Re-cast every matrix in ``params`` with ONE launch (passt_cast_multi).  The pointer table lives on the device
        and is rebuilt only when a parameter or its bf16 copy moved; building it is host work (one small H2D copy),
        so it must first happen outside CUDA-graph capture -- the eager warm-up steps do that."""
        recs, ents = [], []
        for p in params:
            key = (p.data_ptr(), tuple(p.shape))
            ent = self._store.get(key)
            n = p.numel()
            if n % 8 != 0 or (ent is not None and ent[2] is not None):
                self.get(p, False, True)         # odd size or a transposed copy to keep fresh: per-matrix path
                continue
            if ent is None or ent[1].device != p.device:
                wb = torch.empty(p.shape[0], n // p.shape[0], dtype=BF16, device=p.device)
            else:
                wb = ent[1]
            recs.append((p.data_ptr(), wb.data_ptr(), n // 8))
            ents.append((key, p, wb))
        if not recs:
            return
        sig = tuple(recs)
        if getattr(self, "_table_sig", None) != sig:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("passt_b200: the bf16 weight table changed during CUDA-graph capture; run one eager "
                                   "training step before capturing")
            rows, blk = [], 0
            for src, dst, n8 in recs:
                rows.append([src, dst, n8, blk])      # first_block in the low 32 bits of the 4th word, pad = 0
                blk += (n8 + 1023) // 1024
            self._table = torch.tensor(rows, dtype=torch.int64).to(ents[0][1].device)
            self._table_sig, self._table_blocks = sig, blk
        L.call("passt_cast_multi", L.ptr(self._table), len(recs), self._table_blocks, L.stream_ptr())
        for key, p, wb in ents:
            self._store[key] = ((p.data_ptr(), p._version), wb, None)

    def clear(self):
        self._store.clear()
        self._dirty = False
        self.gen += 1
        self._table_sig = None

    def invalidate(self):
        """Force a refresh on next use (parameters were updated without Python seeing it, e.g. by a graph replay)."""
        for k, (ver, wb, wt) in list(self._store.items()):
            self._store[k] = (None, wb, wt)
        self.gen += 1


B_KN = 16           # passt_gemm_bf16 mode flag: B is [K, N] row-major (kBRowMajorKN)
# A/B switches (environment, read once)
#   PASST_B200_FUSE_RESID : residual adds (x + proj(att), x + fc2(act)) run in the proj / fc2 GEMM epilogues (fp32 output
#                           = acc + bias + residual); the LayerNorm pass then only reads the fp32 stream.  Default off:
#                           the fp32 epilogue (twice the store bytes, 16-column chunks) costs what the LN pass saves.
#   PASST_B200_FUSE_DSUM  : attention backward's D = rowsum(dO o O) is accumulated by the proj-dgrad GEMM epilogue
import os as _os
FUSE_RESID = _os.environ.get("PASST_B200_FUSE_RESID", "0") != "0"    # measured: +0.58 ms GEMM epilogue vs -0.48 ms LN
FUSE_DSUM = _os.environ.get("PASST_B200_FUSE_DSUM", "1") != "0"
#   PASST_B200_FUSE_PE    : patch embedding as ONE kernel (TMA patch gather -> smem operand -> tcgen05 GEMM -> token table);
#                           off (default) = passt_im2col (bf16 patch rows in HBM, kept patches only) + the generic GEMM.
#                           The single kernel is correct and tested but slower at the bench shape: its TMA boxes are
#                           16 rows x 80 bytes per patch, and the TMA unit is row-request bound on such short rows.
FUSE_PE = _os.environ.get("PASST_B200_FUSE_PE", "0") != "0"          # measured: 145 us vs 17 + 43 us (80-byte TMA rows)
GEMM_TRACE = None   # bench.py sets this to a list to collect (start_event, end_event, flops) per GEMM launch


def _gemm(A, Bm, C, *, C2=None, bias=None, aux=None, M, N, K, lda, ldb, ldc, mode, period=0, ld_aux=0, splits=1):
    if GEMM_TRACE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    L.call("passt_gemm_bf16", L.ptr(A), L.ptr(Bm), L.ptr(C), L.ptr(C2), L.ptr(bias), L.ptr(aux), M, N, K, lda, ldb,
           ldc, mode, period, ld_aux, splits, 0, L.stream_ptr())
    if GEMM_TRACE is not None:
        e1.record()
        GEMM_TRACE.append((e0, e1, 2.0 * M * N * K))


def _wgrad_splits(M_out: int, N_out: int, tokens: int, clusters: int = 74) -> int:
    """Split-K factor for the weight-gradient GEMM (2-CTA kernel: 256x256 cluster tiles over 74 clusters).
    Minimises rounds x (k-blocks per split + epilogue cost): enough items to fill the chip in whole waves, few enough
    that the fp32 reduce-add epilogue (one full tile per split) stays amortised."""
    tiles = ((M_out + 255) // 256) * (N_out // 256)
    kb = (tokens + 63) // 64
    epi = 8
    best, best_cost = 1, None
    for s in range(1, min(kb, 48) + 1):
        per = (kb + s - 1) // s
        s_eff = (kb + per - 1) // per          # the kernel drops empty splits
        rounds = (tiles * s_eff + clusters - 1) // clusters
        cost = rounds * (per + epi)
        if best_cost is None or cost < best_cost:
            best, best_cost = s_eff, cost
    return best


PARAM_ORDER_HEAD = ["cls_token", "dist_token", "new_pos_embed", "freq_new_pos_embed", "time_new_pos_embed",
                    "patch_embed.proj.weight", "patch_embed.proj.bias"]
PARAM_ORDER_BLOCK = ["norm1.weight", "norm1.bias", "attn.qkv.weight", "attn.qkv.bias", "attn.proj.weight",
                     "attn.proj.bias", "norm2.weight", "norm2.bias", "mlp.fc1.weight", "mlp.fc1.bias",
                     "mlp.fc2.weight", "mlp.fc2.bias"]
PARAM_ORDER_TAIL = ["norm.weight", "norm.bias", "head.0.weight", "head.0.bias", "head.1.weight", "head.1.bias"]


def param_names(depth: int) -> List[str]:
    names = list(PARAM_ORDER_HEAD)
    for i in range(depth):
        names += [f"blocks.{i}.{n}" for n in PARAM_ORDER_BLOCK]
    return names + PARAM_ORDER_TAIL


class SplitWeightCache:
    """fp32-parity tier: [out, 3*in] bf16 operands [W_hi | W_lo | W_hi] of the fp32 master weights, rebuilt when a
    parameter's version changed or the bf16 cache was marked dirty by a training step."""

    def __init__(self):
        self._store: Dict[tuple, tuple] = {}
        self.gen = -1             # WeightCache.gen the split operands were last rebuilt at

    def get(self, p: torch.Tensor, force: bool):
        key = (p.data_ptr(), tuple(p.shape))
        ent = self._store.get(key)
        ver = (p.data_ptr(), p._version)
        if not force and ent is not None and ent[0] == ver:
            return ent[1]
        w2 = p.detach().reshape(p.shape[0], -1)
        R, C = w2.shape
        ws = ent[1] if ent is not None and ent[1].device == p.device else torch.empty(R, 3 * C, dtype=BF16, device=p.device)
        L.call("passt_split3_bf16", L.ptr(w2), L.ptr(ws), R, C, C, 1, L.stream_ptr())
        self._store[key] = (ver, ws)
        return ws


def _split_rows(w2d: torch.Tensor, pattern: int) -> torch.Tensor:
    """fp32 [R, C] -> bf16 [3R, C] row-stacked split (contraction along rows; csrc/fp32tier_bwd.cu)."""
    R, C = w2d.shape
    out = torch.empty(3 * R, C, dtype=BF16, device=w2d.device)
    L.call("passt_split3_rows_bf16", L.ptr(w2d), L.ptr(out), R, C, w2d.stride(0), pattern, L.stream_ptr())
    return out


def _forward_fp32_tier(x32, net, plan: StepPlan, mix, P, save=None):
    """fp32-parity forward (north_star 1e-3 tier): every GEMM runs on the tcgen05 bf16 tensor cores over hi/lo-split
    operands with a 3x longer contraction and fp32 output (csrc/fp32tier.cu), LayerNorm / GELU / attention / residual
    stream in fp32.  save: dict that receives what the fp32-tier backward needs (None under no_grad)."""
    dev = x32.device
    depth, Dm, H = len(net.blocks), net.embed_dim, net.num_heads
    hidden = P["blocks.0.mlp.fc1.weight"].shape[0] if depth else 4 * Dm
    B, ntok = plan.B, plan.ntok
    M = B * ntok
    Fg, Tg = net.patch_embed.grid_size
    fs, ts = net.stride
    st = L.stream_ptr()
    f32 = dict(device=dev, dtype=torch.float32)
    b16 = dict(device=dev, dtype=BF16)
    sw = getattr(net, "_wsplit", None)
    if sw is None:
        sw = net._wsplit = SplitWeightCache()
    refresh = sw.gen != net._wcache.gen      # a training step / graph replay / optimizer pass happened since
    sw.gen = net._wcache.gen

    def gemm_f32(As, Ws, out, tab, period, N, K3):
        _gemm(As, Ws, out, aux=tab, M=M, N=N, K=K3, lda=K3, ldb=K3, ldc=N, mode=2, period=period, ld_aux=N)

    mix_perm, mix_lam = mix if mix is not None else (None, None)
    A0 = torch.empty(M, 256, **f32)
    L.call("passt_im2col_f32", L.ptr(x32), L.ptr(A0), L.ptr(plan.patch_f), L.ptr(plan.patch_t), B, ntok, plan.Fm, plan.Tm,
           fs, ts, L.ptr(mix_perm), L.ptr(mix_lam), st)
    A0s = torch.empty(M, 768, **b16)
    L.call("passt_split3_bf16", L.ptr(A0), L.ptr(A0s), M, 256, 256, 0, st)
    tab = torch.empty(ntok, Dm, **f32)
    L.call("passt_token_table", L.ptr(tab), L.ptr(P["cls_token"]), L.ptr(P["dist_token"]), L.ptr(P["new_pos_embed"]),
           L.ptr(P["patch_embed.proj.bias"]), L.ptr(P["time_new_pos_embed"]), L.ptr(P["freq_new_pos_embed"]),
           L.ptr(plan.patch_f), L.ptr(plan.patch_t), ntok, Fg, Tg, plan.toffset, L.ptr(plan.toffset_dev), st)
    xcur = torch.empty(M, Dm, **f32)
    gemm_f32(A0s, sw.get(P["patch_embed.proj.weight"], refresh), xcur, tab, ntok, Dm, 768)
    delta = None
    scale = float((Dm // H) ** -0.5)
    blocks_saved = []
    for i in range(depth):
        pre = f"blocks.{i}."
        h1 = torch.empty(M, 3 * Dm, **b16)
        x_in = xcur if delta is None else torch.empty(M, Dm, **f32)
        L.call("passt_ln_fwd_f32tier", L.ptr(xcur), L.ptr(delta), None if delta is None else L.ptr(x_in), L.ptr(h1),
               L.ptr(P[pre + "norm1.weight"]), L.ptr(P[pre + "norm1.bias"]), M, Dm, 1e-6, st)
        qkv = torch.empty(M, 3 * Dm, **f32)
        gemm_f32(h1, sw.get(P[pre + "attn.qkv.weight"], refresh), qkv, P[pre + "attn.qkv.bias"], 1, 3 * Dm, 3 * Dm)
        att = torch.empty(M, 3 * Dm, **b16)
        L.call("passt_attn_fwd_f32", L.ptr(qkv), L.ptr(att), B, ntok, H, scale, st)
        oproj = torch.empty(M, Dm, **f32)
        gemm_f32(att, sw.get(P[pre + "attn.proj.weight"], refresh), oproj, P[pre + "attn.proj.bias"], 1, Dm, 3 * Dm)
        x_mid = torch.empty(M, Dm, **f32)
        h2 = torch.empty(M, 3 * Dm, **b16)
        L.call("passt_ln_fwd_f32tier", L.ptr(x_in), L.ptr(oproj), L.ptr(x_mid), L.ptr(h2),
               L.ptr(P[pre + "norm2.weight"]), L.ptr(P[pre + "norm2.bias"]), M, Dm, 1e-6, st)
        pre_act = torch.empty(M, hidden, **f32)
        gemm_f32(h2, sw.get(P[pre + "mlp.fc1.weight"], refresh), pre_act, P[pre + "mlp.fc1.bias"], 1, hidden, 3 * Dm)
        act = torch.empty(M, 3 * hidden, **b16)
        L.call("passt_gelu_split3", L.ptr(pre_act), L.ptr(act), M, hidden, st)
        ofc2 = torch.empty(M, Dm, **f32)
        gemm_f32(act, sw.get(P[pre + "mlp.fc2.weight"], refresh), ofc2, P[pre + "mlp.fc2.bias"], 1, Dm, 3 * hidden)
        if save is not None:
            blocks_saved.append(dict(x_in=x_in, h1=h1, qkv=qkv, att=att, x_mid=x_mid, h2=h2, pre_act=pre_act, act=act))
        xcur, delta = x_mid, ofc2
    if delta is not None:
        x_fin = torch.empty(M, Dm, **f32)
        L.call("passt_ln_fwd_f32tier", L.ptr(xcur), L.ptr(delta), L.ptr(x_fin), None, None, None, M, Dm, 1e-6, st)
        xcur = x_fin
    C = P["head.1.weight"].shape[0]
    logits = torch.empty(B, C, **f32)
    feats = torch.empty(B, Dm, **f32)
    fl = torch.empty(B, Dm, **f32)
    L.call("passt_head_fwd", L.ptr(xcur), None, L.ptr(P["norm.weight"]), L.ptr(P["norm.bias"]),
           L.ptr(P["head.0.weight"]), L.ptr(P["head.0.bias"]), L.ptr(P["head.1.weight"]), L.ptr(P["head.1.bias"]),
           L.ptr(logits), L.ptr(feats), L.ptr(fl), B, ntok, C, st)
    if save is not None:
        save.update(blocks=blocks_saved, A0s=A0s, x_last=xcur, fl=fl, M=M, hidden=hidden, C=C, scale=scale)
    return logits, feats


def _backward_fp32_tier(ctx, dlogits, dfeats):
    """Backward of the fp32 tier: split-operand dgrad / wgrad GEMMs on the tensor cores (contraction 3x longer, fp32
    output), fp32 LayerNorm / GELU / attention backward (csrc/fp32tier_bwd.cu).  Gradients within 1e-3 of the fp32
    reference (tests/test_gpu_parity.py::test_fp32_tier_gradients_1e3)."""
    net, plan, names, params = ctx.net, ctx.plan, ctx.names, ctx.params
    P = dict(zip(names, params))
    mi = ctx.misc
    M, hidden, C, scale = mi["M"], mi["hidden"], mi["C"], mi["scale"]
    depth, Dm, H = len(net.blocks), net.embed_dim, net.num_heads
    B, ntok = plan.B, plan.ntok
    Fg, Tg = net.patch_embed.grid_size
    dev = mi["x_last"].device
    st = L.stream_ptr()
    f32 = dict(device=dev, dtype=torch.float32)
    b16 = dict(device=dev, dtype=BF16)
    sizes = [p.numel() for p in params]
    flat = torch.zeros(sum(sizes), **f32)
    G, off = {}, 0
    for n, p, sz in zip(names, params, sizes):
        G[n] = flat[off: off + sz].view(p.shape)
        off += sz
    ncl = L.load().passt_get_sm_limit() // 2

    def dgrad(dy, w, n_out):
        """dX [M, n_out] = dY [M, K] W [K, n_out] (W = the nn.Linear weight [out=K, in=n_out])."""
        K = dy.shape[1]
        a = torch.empty(M, 3 * K, **b16)
        L.call("passt_split3_bf16", L.ptr(dy), L.ptr(a), M, K, K, 0, st)
        wr = _split_rows(w.detach().reshape(w.shape[0], -1), 1)              # [3K, n_out]
        out = torch.empty(M, n_out, **f32)
        _gemm(a, wr, out, M=M, N=n_out, K=3 * K, lda=3 * K, ldb=n_out, ldc=n_out, mode=2 | B_KN, period=1, ld_aux=n_out)
        return out

    def wgrad(dy, x_split, gw, n_in):
        """dW [K_out, n_in] += dY^T [K_out, M] X [M, n_in]; x_split = the forward's column-stacked operand [M, 3 n_in]."""
        K_out = dy.shape[1]
        a = _split_rows(dy, 0)                                                # [3M, K_out]
        xr = torch.empty(3 * M, n_in, **b16)
        L.call("passt_restack3_bf16", L.ptr(x_split), L.ptr(xr), M, n_in, 1, st)
        _gemm(a, xr, gw, M=K_out, N=n_in, K=3 * M, lda=K_out, ldb=n_in, ldc=n_in, mode=4,
              splits=_wgrad_splits(K_out, n_in, 3 * M, ncl))

    def colsum(dy, gb):
        L.call("passt_colsum_f32", L.ptr(dy), L.ptr(gb), M, dy.shape[1], st)

    dl = torch.zeros(B, C, **f32) if dlogits is None else dlogits.detach().float().contiguous()
    df = None if dfeats is None else dfeats.detach().float().contiguous()
    g = torch.zeros(M, Dm, **f32)
    gb_unused = torch.empty(M, Dm, **b16)
    L.call("passt_head_bwd", L.ptr(mi["x_last"]), None, L.ptr(P["norm.weight"]), L.ptr(P["norm.bias"]),
           L.ptr(P["head.0.weight"]), L.ptr(P["head.0.bias"]), L.ptr(P["head.1.weight"]), L.ptr(dl), L.ptr(df),
           L.ptr(mi["fl"]), L.ptr(g), L.ptr(gb_unused), L.ptr(G["norm.weight"]), L.ptr(G["norm.bias"]),
           L.ptr(G["head.0.weight"]), L.ptr(G["head.0.bias"]), L.ptr(G["head.1.weight"]), L.ptr(G["head.1.bias"]),
           None, B, ntok, C, st)
    ws = torch.empty(3 * B * H * ntok, **f32)
    for i in reversed(range(depth)):
        pre = f"blocks.{i}."
        S = mi["blocks"][i]
        # ---- MLP: x_next = x_mid + fc2(gelu(fc1(LN2(x_mid))))
        colsum(g, G[pre + "mlp.fc2.bias"])
        dact = dgrad(g, P[pre + "mlp.fc2.weight"], hidden)
        wgrad(g, S["act"], G[pre + "mlp.fc2.weight"], hidden)
        dpre = torch.empty(M, hidden, **f32)
        L.call("passt_gelu_bwd_f32", L.ptr(dact), L.ptr(S["pre_act"]), L.ptr(dpre), M * hidden, st)
        colsum(dpre, G[pre + "mlp.fc1.bias"])
        dh2 = dgrad(dpre, P[pre + "mlp.fc1.weight"], Dm)
        wgrad(dpre, S["h2"], G[pre + "mlp.fc1.weight"], Dm)
        g_mid = torch.empty(M, Dm, **f32)
        L.call("passt_ln_bwd_f32", L.ptr(dh2), L.ptr(S["x_mid"]), L.ptr(P[pre + "norm2.weight"]), L.ptr(g), L.ptr(g_mid),
               L.ptr(G[pre + "norm2.weight"]), L.ptr(G[pre + "norm2.bias"]), M, Dm, 1e-6, st)
        # ---- attention: x_mid = x_in + proj(attn(LN1(x_in)))
        colsum(g_mid, G[pre + "attn.proj.bias"])
        datt = dgrad(g_mid, P[pre + "attn.proj.weight"], Dm)
        wgrad(g_mid, S["att"], G[pre + "attn.proj.weight"], Dm)
        dqkv = torch.empty(M, 3 * Dm, **f32)
        L.call("passt_attn_bwd_f32", L.ptr(S["qkv"]), L.ptr(datt), L.ptr(dqkv), L.ptr(ws), B, ntok, H, scale, st)
        colsum(dqkv, G[pre + "attn.qkv.bias"])
        dh1 = dgrad(dqkv, P[pre + "attn.qkv.weight"], Dm)
        wgrad(dqkv, S["h1"], G[pre + "attn.qkv.weight"], Dm)
        g_new = torch.empty(M, Dm, **f32)
        L.call("passt_ln_bwd_f32", L.ptr(dh1), L.ptr(S["x_in"]), L.ptr(P[pre + "norm1.weight"]), L.ptr(g_mid),
               L.ptr(g_new), L.ptr(G[pre + "norm1.weight"]), L.ptr(G[pre + "norm1.bias"]), M, Dm, 1e-6, st)
        g = g_new
        mi["blocks"][i] = None
    # ---- patch embedding
    wgrad(g, mi["A0s"], G["patch_embed.proj.weight"].view(Dm, 256), 256)
    L.call("passt_token_table_bwd", L.ptr(g), L.ptr(G["cls_token"]), L.ptr(G["dist_token"]), L.ptr(G["new_pos_embed"]),
           L.ptr(G["patch_embed.proj.bias"]), L.ptr(G["time_new_pos_embed"]), L.ptr(G["freq_new_pos_embed"]),
           L.ptr(plan.patch_f), L.ptr(plan.patch_t), B, ntok, Fg, Tg, plan.toffset, L.ptr(plan.toffset_dev), st)
    net._last_flat_grad = flat
    ctx.misc = None
    grads = [G[n] if p.requires_grad else None for n, p in zip(names, params)]
    return (None, None, None, None, *grads)


class PasstFunction(torch.autograd.Function):
    """(mel image, *params) -> (logits, features).  Saves bf16 activations for the hand-written backward."""

    @staticmethod
    def forward(ctx, x, net, plan: StepPlan, mix, *params):
        if not x.is_cuda:
            raise RuntimeError("passt_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        dev = x.device
        names = param_names(len(net.blocks))
        P = dict(zip(names, params))
        depth, Dm, H = len(net.blocks), net.embed_dim, net.num_heads
        hidden = P["blocks.0.mlp.fc1.weight"].shape[0] if depth else 4 * Dm
        B, ntok = plan.B, plan.ntok
        M = B * ntok
        Fg, Tg = net.patch_embed.grid_size
        fs, ts = net.stride
        need_grad = plan.grad_mode and any(ctx.needs_input_grad[4:])
        wc: WeightCache = net._wcache
        if need_grad:
            refresh = True          # an optimizer step may have happened since the last forward (see WeightCache)
            wc.dirty = True
        else:
            refresh = wc.dirty
            wc.dirty = False
        st = L.stream_ptr()
        x32 = x.detach()
        if x32.dtype != torch.float32:
            x32 = x32.float()
        x32 = x32.contiguous()
        if x32.shape[1] != 1:
            raise RuntimeError("passt_b200 supports in_channels == 1 (mono spectrograms)")
        if getattr(net, "precision", "bf16") == "fp32":
            if need_grad:
                wc.gen += 1             # parameters may change before the next forward: the split copies must be rebuilt
                save = {}
                out = _forward_fp32_tier(x32, net, plan, mix, P, save)
                ctx.net, ctx.plan, ctx.names, ctx.params = net, plan, names, params
                ctx.misc = dict(save, tier="fp32")
                return out
            wc.dirty = refresh          # the bf16 copies were NOT refreshed by this call: leave their flag as it was
            if refresh:
                wc.gen -= 1             # (re-setting the flag must not count as a new parameter generation)
            return _forward_fp32_tier(x32, net, plan, mix, P)
        f32 = dict(device=dev, dtype=torch.float32)
        b16 = dict(device=dev, dtype=BF16)

        # ---- patch embedding: kept patches -> GEMM + token table (bias + pos embeds + cls/dist rows)
        mix_perm = mix_lam = None
        if mix is not None:
            mix_perm, mix_lam = mix
        # TMA needs 16-byte aligned mel rows (not the 998-frame test shape) and a clip at least one strip box long; the
        # epilogue handles at most one clip boundary per 32 token rows
        use_pe = FUSE_PE and plan.Tm % 4 == 0 and plan.Tm >= 160 and ntok >= 32
        A0 = None
        if not use_pe:
            A0 = torch.empty(M, 256, **b16)
            L.call("passt_im2col", L.ptr(x32), L.ptr(A0), L.ptr(plan.patch_f), L.ptr(plan.patch_t), B, ntok, plan.Fm,
                   plan.Tm, fs, ts, L.ptr(mix_perm), L.ptr(mix_lam), st)
        tab = torch.empty(ntok, Dm, **f32)
        L.call("passt_token_table", L.ptr(tab), L.ptr(P["cls_token"]), L.ptr(P["dist_token"]),
               L.ptr(P["new_pos_embed"]), L.ptr(P["patch_embed.proj.bias"]), L.ptr(P["time_new_pos_embed"]),
               L.ptr(P["freq_new_pos_embed"]), L.ptr(plan.patch_f), L.ptr(plan.patch_t), ntok, Fg, Tg, plan.toffset,
               L.ptr(plan.toffset_dev), st)
        if refresh and wc.fresh_from_optimizer:
            wc.fresh_from_optimizer = False      # FusedAdamW rewrote the copies in its own pass
            refresh = False
        if refresh:
            wnames = ["patch_embed.proj.weight"] + [f"blocks.{i}.{w}.weight" for i in range(depth)
                                                    for w in ("attn.qkv", "attn.proj", "mlp.fc1", "mlp.fc2")]
            wc.refresh_all([P[n] for n in wnames])
        wpe, _ = wc.get(P["patch_embed.proj.weight"], False)
        xcur = torch.empty(M, Dm, **f32)
        if use_pe:
            L.call("passt_patch_embed", L.ptr(x32), L.ptr(wpe), L.ptr(tab), L.ptr(xcur), L.ptr(plan.patch_f),
                   L.ptr(plan.patch_t), B, ntok, plan.Fm, plan.Tm, fs, ts, L.ptr(mix_perm), L.ptr(mix_lam), st)
        else:
            _gemm(A0, wpe, xcur, aux=tab, M=M, N=Dm, K=256, lda=256, ldb=256, ldc=Dm, mode=2, period=ntok, ld_aux=Dm)

        saved = []
        delta = None
        scale = float((Dm // H) ** -0.5)
        fuse = FUSE_RESID
        for i in range(depth):
            pre = f"blocks.{i}."
            wqkv, _ = wc.get(P[pre + "attn.qkv.weight"], False)
            wproj, _ = wc.get(P[pre + "attn.proj.weight"], False)
            wfc1, _ = wc.get(P[pre + "mlp.fc1.weight"], False)
            wfc2, _ = wc.get(P[pre + "mlp.fc2.weight"], False)
            # x_in = xcur (+ delta of the previous block); h1 = LN1(x_in)
            h1 = torch.empty(M, Dm, **b16)
            mean1 = torch.empty(M, **f32); rstd1 = torch.empty(M, **f32)
            if delta is None:
                x_in = xcur
                L.call("passt_ln_fwd", L.ptr(x_in), None, None, L.ptr(h1), L.ptr(mean1), L.ptr(rstd1),
                       L.ptr(P[pre + "norm1.weight"]), L.ptr(P[pre + "norm1.bias"]), M, Dm, 1e-6, st)
            else:
                x_in = torch.empty(M, Dm, **f32)
                L.call("passt_ln_fwd", L.ptr(xcur), L.ptr(delta), L.ptr(x_in), L.ptr(h1), L.ptr(mean1), L.ptr(rstd1),
                       L.ptr(P[pre + "norm1.weight"]), L.ptr(P[pre + "norm1.bias"]), M, Dm, 1e-6, st)
            qkv = torch.empty(M, 3 * Dm, **b16)
            _gemm(h1, wqkv, qkv, bias=P[pre + "attn.qkv.bias"], M=M, N=3 * Dm, K=Dm, lda=Dm, ldb=Dm, ldc=3 * Dm, mode=0)
            att = torch.empty(M, Dm, **b16)
            lse = torch.empty(B, H, ((ntok + 127) // 128) * 128, **f32)   # log2-domain, padded to whole query tiles
            L.call("passt_attn_fwd", L.ptr(qkv), L.ptr(att), L.ptr(lse), B, ntok, H, scale, st)
            x_mid = torch.empty(M, Dm, **f32)
            h2 = torch.empty(M, Dm, **b16)
            mean2 = torch.empty(M, **f32); rstd2 = torch.empty(M, **f32)
            if fuse:
                # x_mid = x_in + att Wproj^T + bias straight from the GEMM epilogue (fp32), then LN reads it once
                _gemm(att, wproj, x_mid, bias=P[pre + "attn.proj.bias"], aux=x_in, M=M, N=Dm, K=Dm, lda=Dm, ldb=Dm,
                      ldc=Dm, mode=2, period=M, ld_aux=Dm)
                L.call("passt_ln_fwd", L.ptr(x_mid), None, None, L.ptr(h2), L.ptr(mean2), L.ptr(rstd2),
                       L.ptr(P[pre + "norm2.weight"]), L.ptr(P[pre + "norm2.bias"]), M, Dm, 1e-6, st)
            else:
                oproj = torch.empty(M, Dm, **b16)
                _gemm(att, wproj, oproj, bias=P[pre + "attn.proj.bias"], M=M, N=Dm, K=Dm, lda=Dm, ldb=Dm, ldc=Dm, mode=0)
                L.call("passt_ln_fwd", L.ptr(x_in), L.ptr(oproj), L.ptr(x_mid), L.ptr(h2), L.ptr(mean2), L.ptr(rstd2),
                       L.ptr(P[pre + "norm2.weight"]), L.ptr(P[pre + "norm2.bias"]), M, Dm, 1e-6, st)
            pre_act = torch.empty(M, hidden, **b16)
            act = torch.empty(M, hidden, **b16)
            _gemm(h2, wfc1, pre_act, C2=act, bias=P[pre + "mlp.fc1.bias"], M=M, N=hidden, K=Dm, lda=Dm, ldb=Dm,
                  ldc=hidden, mode=1)
            if fuse:
                x_next = torch.empty(M, Dm, **f32)
                _gemm(act, wfc2, x_next, bias=P[pre + "mlp.fc2.bias"], aux=x_mid, M=M, N=Dm, K=hidden, lda=hidden,
                      ldb=hidden, ldc=Dm, mode=2, period=M, ld_aux=Dm)
            else:
                ofc2 = torch.empty(M, Dm, **b16)
                _gemm(act, wfc2, ofc2, bias=P[pre + "mlp.fc2.bias"], M=M, N=Dm, K=hidden, lda=hidden, ldb=hidden,
                      ldc=Dm, mode=0)
            if need_grad:
                saved.append(dict(x_in=x_in, mean1=mean1, rstd1=rstd1, h1=h1, qkv=qkv, att=att, lse=lse, x_mid=x_mid,
                                  mean2=mean2, rstd2=rstd2, h2=h2, pre_act=pre_act, act=act))
            if fuse:
                xcur, delta = x_next, None
            else:
                xcur, delta = x_mid, ofc2

        C = P["head.1.weight"].shape[0]
        logits = torch.empty(B, C, **f32)
        feats = torch.empty(B, Dm, **f32)
        fl = torch.empty(B, Dm, **f32)
        L.call("passt_head_fwd", L.ptr(xcur), L.ptr(delta), L.ptr(P["norm.weight"]), L.ptr(P["norm.bias"]),
               L.ptr(P["head.0.weight"]), L.ptr(P["head.0.bias"]), L.ptr(P["head.1.weight"]), L.ptr(P["head.1.bias"]),
               L.ptr(logits), L.ptr(feats), L.ptr(fl), B, ntok, C, st)
        if need_grad:
            ctx.net = net
            ctx.plan = plan
            ctx.names = names
            ctx.params = params
            ctx.saved = saved
            ctx.misc = dict(A0=A0, x32=x32 if A0 is None else None, mix=(mix_perm, mix_lam), x_last=xcur, delta_last=delta, fl=fl, M=M, hidden=hidden, C=C, scale=scale)
        return logits, feats

    @staticmethod
    def backward(ctx, dlogits, dfeats):
        if ctx.misc.get("tier") == "fp32":
            return _backward_fp32_tier(ctx, dlogits, dfeats)
        net, plan, names, params = ctx.net, ctx.plan, ctx.names, ctx.params
        P = dict(zip(names, params))
        mi = ctx.misc
        M, hidden, C, scale = mi["M"], mi["hidden"], mi["C"], mi["scale"]
        depth, Dm, H = len(net.blocks), net.embed_dim, net.num_heads
        B, ntok = plan.B, plan.ntok
        Fg, Tg = net.patch_embed.grid_size
        dev = mi["x_last"].device
        st = L.stream_ptr()
        wc: WeightCache = net._wcache
        f32 = dict(device=dev, dtype=torch.float32)
        b16 = dict(device=dev, dtype=BF16)

        # one flat, zero-initialised fp32 gradient buffer; each parameter's gradient is a view into it
        sizes = [p.numel() for p in params]
        flat = torch.zeros(sum(sizes), **f32)
        G, off, span = {}, 0, {}
        for n, p, s in zip(names, params, sizes):
            G[n] = flat[off: off + s].view(p.shape)
            span[n] = (off, off + s)
            off += s
        hook = getattr(net, "_grad_chunk_hook", None)
        # data parallel: leave a few SMs to the NCCL all-reduce kernels that run next to the backward's persistent kernels
        reserve = int(getattr(getattr(hook, "__self__", None), "reserve_sms", 0) or 0) if hook is not None else 0
        lib = L.load()
        sm_before = lib.passt_get_sm_limit()
        if reserve > 0:
            lib.passt_set_sm_limit(sm_before - reserve)
        ncl = lib.passt_get_sm_limit() // 2

        def chunk_done(first: str, last: str):
            if hook is not None:
                hook(flat, span[first][0], span[last][1])

        dl = None if dlogits is None else dlogits.detach().float().contiguous()
        df = None if dfeats is None else dfeats.detach().float().contiguous()
        if dl is None:
            dl = torch.zeros(B, C, **f32)
        g = torch.zeros(M, Dm, **f32)
        gb = torch.zeros(M, Dm, **b16)
        last_bias = G[f"blocks.{depth - 1}.mlp.fc2.bias"] if depth else None
        L.call("passt_head_bwd", L.ptr(mi["x_last"]), L.ptr(mi["delta_last"]), L.ptr(P["norm.weight"]),
               L.ptr(P["norm.bias"]), L.ptr(P["head.0.weight"]), L.ptr(P["head.0.bias"]), L.ptr(P["head.1.weight"]),
               L.ptr(dl), L.ptr(df), L.ptr(mi["fl"]), L.ptr(g), L.ptr(gb), L.ptr(G["norm.weight"]), L.ptr(G["norm.bias"]),
               L.ptr(G["head.0.weight"]), L.ptr(G["head.0.bias"]), L.ptr(G["head.1.weight"]), L.ptr(G["head.1.bias"]),
               L.ptr(last_bias), B, ntok, C, st)
        chunk_done("norm.weight", "head.1.bias")

        ws_bytes = L.load().passt_attn_bwd_workspace_bytes(B, ntok, H)
        attn_ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
        dsum_ptr = L.load().passt_attn_bwd_dsum_ptr(L.ptr(attn_ws), B, ntok, H) if FUSE_DSUM else None
        dact = torch.empty(M, hidden, **b16)
        dh = torch.empty(M, Dm, **b16)
        datt = torch.empty(M, Dm, **b16)
        dqkv = torch.empty(M, 3 * Dm, **b16)
        for i in reversed(range(depth)):
            pre = f"blocks.{i}."
            S = ctx.saved[i]
            # dgrad GEMMs read the bf16 weight itself as a [K, N] row-major (MN-major) B operand: no transposed copies
            wqkv, _ = wc.get(P[pre + "attn.qkv.weight"], False)
            wproj, _ = wc.get(P[pre + "attn.proj.weight"], False)
            wfc1, _ = wc.get(P[pre + "mlp.fc1.weight"], False)
            wfc2, _ = wc.get(P[pre + "mlp.fc2.weight"], False)
            # ---- MLP
            _gemm(gb, wfc2, dact, aux=S["pre_act"], bias=G[pre + "mlp.fc1.bias"], M=M, N=hidden, K=Dm, lda=Dm,
                  ldb=hidden, ldc=hidden, mode=3 | B_KN, ld_aux=hidden)   # d pre = (g W2) * gelu'(pre); + fc1 bias grad
            _gemm(gb, S["act"], G[pre + "mlp.fc2.weight"], M=Dm, N=hidden, K=M, lda=Dm, ldb=hidden, ldc=hidden,
                  mode=4, splits=_wgrad_splits(Dm, hidden, M, ncl))
            _gemm(dact, wfc1, dh, M=M, N=Dm, K=hidden, lda=hidden, ldb=Dm, ldc=Dm, mode=0 | B_KN)
            _gemm(dact, S["h2"], G[pre + "mlp.fc1.weight"], M=hidden, N=Dm, K=M, lda=hidden, ldb=Dm, ldc=Dm, mode=4,
                  splits=_wgrad_splits(hidden, Dm, M, ncl))
            L.call("passt_ln_bwd", L.ptr(dh), L.ptr(S["x_mid"]), L.ptr(S["mean2"]), L.ptr(S["rstd2"]),
                   L.ptr(P[pre + "norm2.weight"]), L.ptr(g), L.ptr(g), L.ptr(gb), L.ptr(G[pre + "norm2.weight"]),
                   L.ptr(G[pre + "norm2.bias"]), L.ptr(G[pre + "attn.proj.bias"]), M, Dm, st)
            # ---- attention
            if FUSE_DSUM:
                # d att = g Wproj with D = rowsum(d att o att) accumulated by the same epilogue (workspace zeroed first)
                L.call("passt_attn_bwd_prepare", L.ptr(attn_ws), B, ntok, H, st)
                _gemm(gb, wproj, datt, bias=dsum_ptr, aux=S["att"], M=M, N=Dm, K=Dm, lda=Dm, ldb=Dm, ldc=Dm,
                      mode=5 | B_KN, period=ntok, ld_aux=Dm)
            else:
                _gemm(gb, wproj, datt, M=M, N=Dm, K=Dm, lda=Dm, ldb=Dm, ldc=Dm, mode=0 | B_KN)
            _gemm(gb, S["att"], G[pre + "attn.proj.weight"], M=Dm, N=Dm, K=M, lda=Dm, ldb=Dm, ldc=Dm, mode=4,
                  splits=_wgrad_splits(Dm, Dm, M, ncl))
            L.call("passt_attn_bwd_ex", L.ptr(S["qkv"]), L.ptr(S["att"]), L.ptr(datt), L.ptr(S["lse"]), L.ptr(dqkv),
                   L.ptr(G[pre + "attn.qkv.bias"]), L.ptr(attn_ws), B, ntok, H, scale, 1 if FUSE_DSUM else 0,
                   st)   # + qkv bias gradient
            _gemm(dqkv, wqkv, dh, M=M, N=Dm, K=3 * Dm, lda=3 * Dm, ldb=Dm, ldc=Dm, mode=0 | B_KN)
            _gemm(dqkv, S["h1"], G[pre + "attn.qkv.weight"], M=3 * Dm, N=Dm, K=M, lda=3 * Dm, ldb=Dm, ldc=Dm, mode=4,
                  splits=_wgrad_splits(3 * Dm, Dm, M, ncl))
            prev_bias = G[f"blocks.{i - 1}.mlp.fc2.bias"] if i > 0 else None
            L.call("passt_ln_bwd", L.ptr(dh), L.ptr(S["x_in"]), L.ptr(S["mean1"]), L.ptr(S["rstd1"]),
                   L.ptr(P[pre + "norm1.weight"]), L.ptr(g), L.ptr(g), L.ptr(gb), L.ptr(G[pre + "norm1.weight"]),
                   L.ptr(G[pre + "norm1.bias"]), L.ptr(prev_bias), M, Dm, st)
            ctx.saved[i] = None
            # every gradient of block i is final except mlp.fc2.bias of block i-1 (written above into ITS span)
            chunk_done(pre + "norm1.weight", pre + "mlp.fc2.bias")
        # ---- patch embedding
        gpe = G["patch_embed.proj.weight"].view(Dm, 256)
        A0 = mi["A0"]
        if A0 is None:
            # the forward never wrote the patch rows to HBM (single-kernel patch embedding): gather them now for the
            # weight-gradient GEMM (kept patches only, same mixup)
            A0 = torch.empty(M, 256, **b16)
            fs, ts = net.stride
            L.call("passt_im2col", L.ptr(mi["x32"]), L.ptr(A0), L.ptr(plan.patch_f), L.ptr(plan.patch_t), B, ntok,
                   plan.Fm, plan.Tm, fs, ts, L.ptr(mi["mix"][0]), L.ptr(mi["mix"][1]), st)
        _gemm(gb, A0, gpe, M=Dm, N=256, K=M, lda=Dm, ldb=256, ldc=256, mode=4, splits=_wgrad_splits(Dm, 256, M, ncl))
        L.call("passt_token_table_bwd", L.ptr(g), L.ptr(G["cls_token"]), L.ptr(G["dist_token"]),
               L.ptr(G["new_pos_embed"]), L.ptr(G["patch_embed.proj.bias"]), L.ptr(G["time_new_pos_embed"]),
               L.ptr(G["freq_new_pos_embed"]), L.ptr(plan.patch_f), L.ptr(plan.patch_t), B, ntok, Fg, Tg,
               plan.toffset, L.ptr(plan.toffset_dev), st)
        chunk_done("cls_token", "patch_embed.proj.bias")
        net._last_flat_grad = flat
        lib.passt_set_sm_limit(sm_before)
        ctx.saved = None
        ctx.misc = None
        grads = [G[n] if p.requires_grad else None for n, p in zip(names, params)]
        return (None, None, None, None, *grads)
