"""PaSST network — drop-in module surface of the reference ``models/passt.py`` on the sm_100a engine.

What is kept from the reference (because callers depend on it):
  * ``get_model(arch, pretrained, n_classes, in_channels, fstride, tstride, input_fdim, input_tdim, u_patchout,
    s_patchout_t, s_patchout_f)`` with the same keywords/defaults and the same 14 arch strings
    (models/passt.py:957-1018), ``lighten_model`` (:932-954), ``fix_embedding_layer`` (:922-929),
    ``get_ensemble_model`` / ``EnsembelerModel`` (:1021-1045);
  * ``PaSST`` as an ``nn.Module`` with real ``nn.Parameter``s under the reference's state_dict keys
    (SURVEY.md §8 a7), ``forward(x[B,1,F,T]) -> (logits, features)`` (:576-595), ``.train()/.eval()`` semantics
    (patchout + random time-embedding offset only in training), attributes ``blocks``, ``patch_embed.grid_size``,
    ``patch_embed.proj``, ``num_tokens``, ``default_cfg``, ``no_weight_decay()``, ``get_classifier()``,
    ``reset_classifier()`` (:486-504).
What is different: ``forward`` does not run torch ops — it hands the parameters to
``passt_b200.engine.PasstFunction`` which launches the hand-written CUDA kernels and a hand-written backward.
The sub-modules (``Block``, ``Attention``, ``Mlp``, ``PatchEmbed``) are parameter containers; calling them
individually is not supported (they raise), so nothing can silently fall back to eager PyTorch.
"""
from __future__ import annotations

import math
import os
import warnings
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import engine

# arch string -> (registry key used for checkpoint lookup, depth, stride the checkpoint was trained with)
ARCHS: Dict[str, Tuple[str, int, Optional[Tuple[int, int]]]] = {
    "passt_deit_bd_p16_384": ("deit_base_distilled_patch16_384", 12, None),
    "passt_s_kd_p16_128_ap486": ("passt_s_kd_p16_128_ap486", 12, (10, 10)),
    "passt_l_kd_p16_128_ap47": ("passt_l_kd_p16_128_ap47", 7, (10, 10)),
    "passt_s_swa_p16_128_ap476": ("passt_s_swa_p16_128_ap476", 12, (10, 10)),
    "passt_s_swa_p16_128_ap4761": ("passt_s_swa_p16_128_ap4761", 12, (10, 10)),
    "passt_s_p16_128_ap472": ("passt_s_p16_128_ap472", 12, (10, 10)),
    "passt_s_p16_s16_128_ap468": ("passt_s_p16_s16_128_ap468", 12, (16, 16)),
    "passt_s_swa_p16_s16_128_ap473": ("passt_s_swa_p16_s16_128_ap473", 12, (16, 16)),
    "passt_s_swa_p16_s14_128_ap471": ("passt_s_swa_p16_s14_128_ap471", 12, (14, 14)),
    "passt_s_p16_s14_128_ap469": ("passt_s_p16_s14_128_ap469", 12, (14, 14)),
    "passt_s_swa_p16_s12_128_ap473": ("passt_s_swa_p16_s12_128_ap473", 12, (12, 12)),
    "passt_s_p16_s12_128_ap470": ("passt_s_p16_s12_128_ap470", 12, (12, 12)),
    "passt_s_f128_20sec_p16_s10_ap474": ("passt-s-f128-20sec-p16-s10-ap474-swa", 12, None),
    "passt_s_f128_30sec_p16_s10_ap473": ("passt-s-f128-30sec-p16-s10-ap473-swa", 12, None),
}
# released checkpoint file names (models/passt.py:175-234); resolved inside $PASST_B200_CKPT_DIR (no network here)
CKPT_FILES = {
    "deit_base_distilled_patch16_384": "deit_base_distilled_patch16_384-d0272ac0.pth",
    "passt_s_swa_p16_128_ap476": "passt-s-f128-p16-s10-ap.476-swa.pt",
    "passt_s_kd_p16_128_ap486": "passt-s-kd-ap.486.pt",
    "passt_l_kd_p16_128_ap47": "passt-l-kd-ap.47.pt",
    "passt_s_swa_p16_128_ap4761": "passt-s-f128-p16-s10-ap.4761-swa.pt",
    "passt_s_p16_128_ap472": "passt-s-f128-p16-s10-ap.472.pt",
    "passt_s_p16_s16_128_ap468": "passt-s-f128-p16-s16-ap.468.pt",
    "passt_s_swa_p16_s16_128_ap473": "passt-s-f128-p16-s16-ap.473-swa.pt",
    "passt_s_swa_p16_s14_128_ap471": "passt-s-f128-p16-s14-ap.471-swa.pt",
    "passt_s_p16_s14_128_ap469": "passt-s-f128-p16-s14-ap.469.pt",
    "passt_s_swa_p16_s12_128_ap473": "passt-s-f128-p16-s12-ap.473-swa.pt",
    "passt_s_p16_s12_128_ap470": "passt-s-f128-p16-s12-ap.470.pt",
    "passt-s-f128-20sec-p16-s10-ap474-swa": "passt-s-f128-20sec-p16-s10-ap.474-swa.pt",
    "passt-s-f128-30sec-p16-s10-ap473-swa": "passt-s-f128-30sec-p16-s10-ap.473-swa.pt",
}


def _trunc_normal_(t: torch.Tensor, std: float = 0.02):
    # reference init: trunc_normal_(std=.02) with cut-offs a=-2, b=2 (vit_helpers.py:277-294)
    return nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2.0, b=2.0)


class _ParamHolder(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError(f"{type(self).__name__} is a parameter container in passt_b200; run the whole PaSST module "
                           "(its forward launches the fused sm_100a kernels)")


class Mlp(_ParamHolder):
    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden_features, in_features)
        self.drop = nn.Dropout(0.0)


class Attention(_ParamHolder):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.attn_drop = nn.Dropout(0.0)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(0.0)


class Block(_ParamHolder):
    def __init__(self, dim, num_heads, mlp_ratio=4.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = Attention(dim, num_heads)
        self.drop_path = nn.Identity()
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))


class PatchEmbed(_ParamHolder):
    def __init__(self, img_size, patch_size, stride, in_chans, embed_dim):
        super().__init__()
        self.img_size = tuple(img_size)
        self.patch_size = (patch_size, patch_size)
        self.stride = tuple(stride)
        self.grid_size = (img_size[0] // stride[0], img_size[1] // stride[1])      # models/passt.py:311
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.flatten = False
        self.embed_dim = embed_dim
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=stride)
        self.norm = nn.Identity()


class PaSST(nn.Module):
    def __init__(self, u_patchout=0, s_patchout_t=0, s_patchout_f=0, img_size=(128, 998), patch_size=16, stride=16,
                 in_chans=1, num_classes=527, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0, qkv_bias=True,
                 distilled=True, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0, **unused):
        super().__init__()
        if not distilled:
            raise NotImplementedError("every PaSST arch is distilled (cls + dist tokens)")
        if drop_rate or attn_drop_rate or drop_path_rate:
            raise NotImplementedError("dropout / stochastic depth are 0 in every PaSST arch (models/passt.py:394)")
        if embed_dim != 768 or embed_dim // num_heads != 64 or patch_size != 16 or in_chans != 1 or not qkv_bias:
            raise NotImplementedError("kernels are built for embed_dim=768, head_dim=64, patch 16, mono input")
        stride = (stride, stride) if isinstance(stride, int) else tuple(stride)
        self.num_classes = num_classes
        self.u_patchout = u_patchout
        self.s_patchout_t = s_patchout_t
        self.s_patchout_f = s_patchout_f
        self.num_features = self.embed_dim = embed_dim
        self.num_tokens = 2
        self.num_heads = num_heads
        self.patch_size = patch_size
        self.stride = stride
        self.patch_embed = PatchEmbed(img_size, patch_size, stride, in_chans, embed_dim)
        Fg, Tg = self.patch_embed.grid_size
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.dist_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.new_pos_embed = nn.Parameter(torch.zeros(1, 2, embed_dim))
        self.freq_new_pos_embed = nn.Parameter(torch.zeros(1, embed_dim, Fg, 1))
        self.time_new_pos_embed = nn.Parameter(torch.zeros(1, embed_dim, 1, Tg))
        self.pos_drop = nn.Dropout(p=0.0)
        self.blocks = nn.Sequential(*[Block(embed_dim, num_heads, mlp_ratio) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        self.pre_logits = nn.Identity()
        self.head = nn.Sequential(nn.LayerNorm(embed_dim),
                                  nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity())
        self.head_dist = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        self.default_cfg = {}
        # arithmetic tier of forward(): "bf16" (tensor-core operands rounded to bf16, 1e-2 parity, training + inference)
        # or "fp32" (hi/lo-split tensor-core GEMMs + fp32 attention, 1e-3 parity, forward-only)
        self.precision = "bf16"
        self._wcache = engine.WeightCache()
        self._mix = None          # optional (perm[B] int32, lam[B] f32) set by fused_mixup()
        self._preset_plan = None  # optional StepPlan prepared by the caller (CUDA-graph replays)
        self.last_plan = None     # StepPlan of the most recent forward (parity tests read the indices)
        self.init_weights()

    # ---- initialisation (models/passt.py:471-484, :598-630) -----------------------------------------------
    def init_weights(self, mode=""):
        for t in (self.new_pos_embed, self.freq_new_pos_embed, self.time_new_pos_embed, self.dist_token,
                  self.cls_token):
            _trunc_normal_(t)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                _trunc_normal_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.LayerNorm):
                nn.init.zeros_(m.bias)
                nn.init.ones_(m.weight)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {"new_pos_embed", "freq_new_pos_embed", "time_new_pos_embed", "cls_token", "dist_token"}

    def get_classifier(self):
        return self.head, self.head_dist

    def reset_classifier(self, num_classes, global_pool=""):
        # same (LayerNorm-dropping) behaviour as the reference (models/passt.py:499-504) is NOT reproduced: the fused
        # head kernel needs head = Sequential(LayerNorm, Linear); we keep the LayerNorm and swap the Linear.
        self.num_classes = num_classes
        dev = self.head[0].weight.device
        self.head = nn.Sequential(self.head[0], nn.Linear(self.embed_dim, num_classes).to(dev))
        self.head_dist = nn.Linear(self.embed_dim, num_classes).to(dev)

    def __deepcopy__(self, memo):
        # SWA deep-copies the net (helpers/swa_callback.py:140): the bf16 weight cache must not be shared
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k == "_wcache":
                new.__dict__[k] = engine.WeightCache()
            elif k == "_wsplit":
                continue
            elif k in ("last_plan", "_mix", "_preset_plan"):
                new.__dict__[k] = None
            else:
                new.__dict__[k] = copy.deepcopy(v, memo)
        return new

    def _ordered_params(self):
        sd = dict(self.named_parameters())
        depth = len(self.blocks)
        # nn.Sequential re-indexes blocks after lighten_model, so names are always blocks.0..depth-1
        return [sd[n] for n in engine.param_names(depth)]

    def fused_mixup(self, perm: Optional[torch.Tensor], lam: Optional[torch.Tensor]):
        """Fold spectrogram mixup x*lam + x[perm]*(1-lam) (ex_audioset.py:173-177) into the patch gather of the
        next forward call.  perm: int [B], lam: float [B] (device tensors)."""
        if perm is None:
            self._mix = None
        else:
            self._mix = (perm.to(torch.int32).contiguous(), lam.to(torch.float32).contiguous())

    # The whole network is ONE opaque autograd node (engine.PasstFunction) around C-ABI launches, preceded by host-side
    # RNG draws: there is nothing for a tracing compiler to fuse.  ``torch.compile(self.net)`` -- the reference's default
    # (compile=True, ex_audioset.py:79,132-135) -- must therefore see the forward as an opaque call: dynamo skips it and
    # runs it eagerly, gradients flow through the same autograd node (tests/test_gpu_boundary.py).
    @torch.compiler.disable
    def forward(self, x):
        if x.dim() != 4:
            raise ValueError(f"expected [B, 1, F, T], got {tuple(x.shape)}")
        if not x.is_cuda:
            raise RuntimeError("passt_b200.PaSST runs on CUDA (sm_100a) only; there is no CPU path")
        Tg = self.patch_embed.grid_size[1]
        t_dim = (x.shape[-1] - self.patch_size) // self.stride[1] + 1
        if not (x.shape[2] == self.patch_embed.img_size[0] and x.shape[3] == self.patch_embed.img_size[1]):
            warnings.warn(f"Input image size ({x.shape[2]}*{x.shape[3]}) doesn't match model "
                          f"({self.patch_embed.img_size[0]}*{self.patch_embed.img_size[1]}).")
        if t_dim >= Tg and t_dim > Tg:
            warnings.warn(f"the patches time dim {t_dim} is larger than the expected time encodings {Tg}, x will be cut")
        with torch.cuda.device(x.device):
            plan = self._preset_plan if self._preset_plan is not None else engine.draw_step_plan(self, x, self.training)
            plan.grad_mode = torch.is_grad_enabled()
            self.last_plan = plan
            if x.shape[0] == 0:
                # empty batch: the reference's ops return empty tensors (the random draws above were still consumed)
                self._mix = None
                n_cls = self.head[1].out_features
                return x.new_zeros(0, n_cls, dtype=torch.float32), x.new_zeros(0, self.embed_dim, dtype=torch.float32)
            mix, self._mix = self._mix, None
            logits, feats = engine.PasstFunction.apply(x, self, plan, mix, *self._ordered_params())
        return logits, feats


# ---- checkpoint adaptation (models/passt.py:656-706, vit_helpers.py:27-51, :54-141) ---------------------------
def adapt_image_pos_embed_to_passt(posemb, num_tokens=1, gs_new=(), mode="bicubic"):
    tok, grid = posemb[:, :num_tokens], posemb[0, num_tokens:]
    gs_old = int(math.sqrt(len(grid)))
    grid = grid.reshape(1, gs_old, gs_old, -1).permute(0, 3, 1, 2)
    grid = F.interpolate(grid, size=gs_new, mode=mode, align_corners=False)
    return tok, grid.mean(dim=3, keepdim=True), grid.mean(dim=2, keepdim=True)


def checkpoint_filter_fn(state_dict, model):
    if "model" in state_dict:
        state_dict = state_dict["model"]
    state_dict = dict(state_dict)
    if "time_new_pos_embed" not in state_dict:
        tok, fpos, tpos = adapt_image_pos_embed_to_passt(state_dict.pop("pos_embed"), model.num_tokens,
                                                         model.patch_embed.grid_size)
        state_dict.update(new_pos_embed=tok, freq_new_pos_embed=fpos, time_new_pos_embed=tpos)
    out = {}
    for k, v in state_dict.items():
        if "patch_embed.proj.weight" in k and v.dim() < 4:
            O, I, Hh, Ww = model.patch_embed.proj.weight.shape
            v = v.reshape(O, -1, Hh, Ww)
        out[k] = v
    return out


def _adapt_input_conv(in_chans, w):
    if in_chans == 1 and w.shape[1] == 3:
        return w.float().sum(dim=1, keepdim=True).to(w.dtype)
    return w


def load_pretrained(model: PaSST, registry_key: str, n_classes: int, pretrained_classes: int = 527):
    ckpt_dir = os.environ.get("PASST_B200_CKPT_DIR", "")
    fname = CKPT_FILES.get(registry_key)
    path = os.path.join(ckpt_dir, fname) if ckpt_dir and fname else None
    if not path or not os.path.isfile(path):
        raise RuntimeError(
            f"pretrained=True needs the released checkpoint '{fname}' in $PASST_B200_CKPT_DIR (this build has no "
            "network access; the reference would download it, models/helpers/vit_helpers.py:85-91). "
            "Pass pretrained=False for random initialisation.")
    sd = checkpoint_filter_fn(torch.load(path, map_location="cpu"), model)
    sd["patch_embed.proj.weight"] = _adapt_input_conv(1, sd["patch_embed.proj.weight"])
    strict = True
    is_deit = registry_key.startswith("deit")
    if is_deit or n_classes != pretrained_classes:
        for k in ("head.1.weight", "head.1.bias", "head_dist.weight", "head_dist.bias", "head.weight", "head.bias"):
            sd.pop(k, None)
        strict = False
    model.load_state_dict(sd, strict=strict)


# ---- factories --------------------------------------------------------------------------------------------------
def fix_embedding_layer(model, embed="default"):
    if embed != "default":
        raise NotImplementedError("only embed='default' exists (the reference's other branches name undefined classes, "
                                  "models/passt.py:926-929)")
    return model


def lighten_model(model, cut_depth=0):
    """Remove transformer blocks (models/passt.py:932-954)."""
    if cut_depth == 0:
        return model
    old = list(model.blocks.children())
    if cut_depth < 0:
        old = [old[0]] + old[1:-1:-cut_depth] + [old[-1]]
    else:
        if len(model.blocks) < cut_depth + 2:
            raise ValueError(f"Cut depth a VIT with {len(model.blocks)} layers should be between 1 and "
                             f"{len(model.blocks) - 2}")
        old = [old[0]] + old[cut_depth + 1:]
    model.blocks = nn.Sequential(*old)
    return model


def get_model(arch="passt_s_kd_p16_128_ap486", pretrained=True, n_classes=527, in_channels=1, fstride=10, tstride=10,
              input_fdim=128, input_tdim=998, u_patchout=0, s_patchout_t=0, s_patchout_f=0):
    """Same signature and defaults as the reference factory (models/passt.py:957-961)."""
    if arch not in ARCHS:
        raise RuntimeError(f"Unknown model {arch}")
    key, depth, trained_stride = ARCHS[arch]
    if trained_stride is not None and (fstride, tstride) != trained_stride:
        warnings.warn(f"This model was pre-trained with strides {trained_stride}, but now you set "
                      f"(fstride,tstride) to {(fstride, tstride)}.")
    model = PaSST(u_patchout=u_patchout, s_patchout_t=s_patchout_t, s_patchout_f=s_patchout_f,
                  img_size=(input_fdim, input_tdim), patch_size=16, stride=(fstride, tstride), in_chans=in_channels,
                  num_classes=n_classes, embed_dim=768, depth=depth, num_heads=12, distilled=True)
    model.default_cfg = {"architecture": key, "num_classes": 527, "input_size": (1, input_fdim, input_tdim),
                         "first_conv": "patch_embed.proj", "classifier": ("head.1", "head_dist")}
    if pretrained:
        load_pretrained(model, key, n_classes)
    model = fix_embedding_layer(model)
    model = lighten_model(model)
    return model


class EnsembelerModel(nn.Module):
    """Average of several nets' logits (models/passt.py:1021-1036).  The nets run on the same spectrogram; the logit
    average is one fused launch (passt_ens_sigmoid) instead of a chain of adds and a divide."""

    def __init__(self, models):
        super().__init__()
        self.models = nn.ModuleList(models)

    def forward(self, x):
        from .evalpath import _sigmoid_mean
        outs = [m(x)[0] for m in self.models]
        _, total = _sigmoid_mean(outs, mode=0, want_mean_logits=True)
        return total, total


def get_ensemble_model(arch_list=[]):
    return EnsembelerModel([get_model(arch=a, fstride=f, tstride=t) for a, f, t in arch_list])
