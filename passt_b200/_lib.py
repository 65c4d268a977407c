"""ctypes binding of libpasst_b200.so (the C ABI declared in include/passt_b200.h).

There is deliberately no CPU / eager fallback: if the library is missing or a kernel fails, we raise.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "lib", "libpasst_b200.so")
_lock = threading.Lock()
_lib = None

vp, i32, f32, f64 = C.c_void_p, C.c_int, C.c_float, C.c_double

_SIGS = {
    "passt_gemm_debug_desc": (None, [i32, vp]),
    "passt_gemm_set_2cta": (None, [i32]),
    "passt_gemm_bf16": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp]),
    "passt_mel_workspace_bytes": (C.c_size_t, []),
    "passt_mel_init": (i32, [vp, i32, vp]),
    "passt_mel_set_band": (i32, [vp, f64, f64, i32, vp]),
    "passt_mel_set_band_dev": (i32, [vp, vp, i32, vp]),
    "passt_mel_forward": (i32, [vp, vp, vp, i32, i32, i32, vp, i32, i32, vp]),
    "passt_ln_fwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, f32, vp]),
    "passt_ln_bwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, vp]),
    "passt_colsum_bf16": (i32, [vp, vp, i32, i32, i32, vp]),
    "passt_im2col": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp]),
    "passt_patch_embed": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp]),
    "passt_token_table": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, vp]),
    "passt_token_table_bwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp]),
    "passt_cast_transpose": (i32, [vp, vp, vp, i32, i32, vp]),
    "passt_cast_multi": (i32, [vp, i32, i32, vp]),
    "passt_adamw_step": (i32, [vp, i32, i32, vp, vp]),
    "passt_head_fwd": (i32, [vp] * 11 + [i32, i32, i32, vp]),
    "passt_head_bwd": (i32, [vp] * 19 + [i32, i32, i32, vp]),
    "passt_attn_fwd": (i32, [vp, vp, vp, i32, i32, i32, f32, vp]),
    "passt_attn_bwd": (i32, [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, f32, vp]),
    "passt_attn_bwd_workspace_bytes": (C.c_size_t, [i32, i32, i32]),
    "passt_loss_workspace_bytes": (C.c_size_t, [i32]),
    "passt_loss_bce": (i32, [vp, vp, vp, vp, vp, vp, vp, i32, i32, vp]),
    "passt_loss_ce": (i32, [vp, vp, vp, vp, vp, vp, vp, i32, i32, vp]),
    "passt_scale_dev": (i32, [vp, vp, vp, C.c_size_t, vp]),
    "passt_ens_sigmoid": (i32, [vp, i32, vp, vp, C.c_size_t, i32, vp]),
    "passt_average_precision": (i32, [vp, vp, vp, i32, i32, vp]),
    "passt_swa_update": (i32, [vp, i32, i32, C.c_longlong, vp]),
    "passt_wave_augment": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp]),
    "passt_im2col_f32": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp]),
    "passt_split3_rows_bf16": (i32, [vp, vp, C.c_longlong, i32, i32, i32, vp]),
    "passt_restack3_bf16": (i32, [vp, vp, C.c_longlong, i32, i32, vp]),
    "passt_colsum_f32": (i32, [vp, vp, C.c_longlong, i32, vp]),
    "passt_ln_apply_f32": (i32, [vp, vp, vp, vp, i32, i32, f32, vp]),
    "passt_ln_bwd_f32": (i32, [vp, vp, vp, vp, vp, vp, vp, i32, i32, f32, vp]),
    "passt_gelu_bwd_f32": (i32, [vp, vp, vp, C.c_longlong, vp]),
    "passt_attn_bwd_f32": (i32, [vp, vp, vp, vp, i32, i32, i32, f32, vp]),
    "passt_split3_bf16": (i32, [vp, vp, C.c_longlong, i32, i32, i32, vp]),
    "passt_gelu_split3": (i32, [vp, vp, C.c_longlong, i32, vp]),
    "passt_ln_fwd_f32tier": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, f32, vp]),
    "passt_attn_fwd_f32": (i32, [vp, vp, i32, i32, i32, f32, vp]),
    "passt_set_sm_limit": (None, [i32]),
    "passt_get_sm_limit": (i32, []),
    "passt_set_pdl": (None, [i32]),
    "passt_get_pdl": (i32, []),
    "passt_attn_fwd_set_variant": (None, [i32]),
    "passt_attn_bwd_set_variant": (None, [i32]),
    "passt_attn_bwd_get_variant": (i32, []),
    "passt_attn_bwd_prepare": (i32, [vp, i32, i32, i32, vp]),
    "passt_attn_bwd_dsum_ptr": (vp, [vp, i32, i32, i32]),
    "passt_attn_bwd_ex": (i32, [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, f32, i32, vp]),
    "passt_attn_debug_timeline": (None, [vp]),
    "passt_attn_bwd_debug_timeline": (None, [vp]),
}


class PasstLibError(RuntimeError):
    pass


def lib_path() -> str:
    return _LIB_PATH


def load():
    """Load the library.  build.build() decides from the source digest (lib/build.sha256) whether the .so has to be
    (re)compiled, so edited .cu sources can never run a stale library; raises if it cannot be built or loaded."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        from . import build as _build
        try:
            _build.build()
        except Exception:
            # sources changed (or no .so) and it cannot be rebuilt here: fail loudly rather than run a stale library,
            # unless the caller explicitly accepts the existing file
            if not (os.path.isfile(_LIB_PATH) and os.environ.get("PASST_B200_ALLOW_STALE_LIB") == "1"):
                raise
        lib = C.CDLL(_LIB_PATH)
        for name, (res, args) in _SIGS.items():
            try:
                fn = getattr(lib, name)
            except AttributeError:
                continue
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def exported_symbols():
    return list(_SIGS)


def ptr(t):
    if t is None:
        return None
    if isinstance(t, int):
        return t
    return t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def check(rc: int, what: str):
    if rc != 0:
        msg = f"{what} failed with code {rc}"
        if rc > 0:
            msg += " (cudaError)"
        raise PasstLibError(msg)


# kernels launched per C-ABI call (for bench.py's gpu_launches claim)
_LAUNCHES = {"passt_adamw_step": 2, "passt_attn_debug_timeline": 0, "passt_attn_bwd_debug_timeline": 0, "passt_gemm_set_2cta": 0, "passt_gemm_debug_desc": 0, "passt_head_bwd": 2, "passt_attn_bwd": 3, "passt_attn_bwd_ex": 3, "passt_attn_bwd_prepare": 0, "passt_attn_bwd_dsum_ptr": 0, "passt_set_pdl": 0, "passt_set_sm_limit": 0, "passt_get_sm_limit": 0, "passt_get_pdl": 0, "passt_attn_fwd_set_variant": 0, "passt_attn_bwd_set_variant": 0, "passt_attn_bwd_get_variant": 0, "passt_mel_workspace_bytes": 0, "passt_loss_workspace_bytes": 0,
             "passt_attn_bwd_workspace_bytes": 0}
_launch_counter = 0


def reset_launch_count():
    global _launch_counter
    _launch_counter = 0


def launch_count() -> int:
    return _launch_counter


def call(name: str, *args):
    global _launch_counter
    fn = getattr(load(), name)
    rc = fn(*args)
    check(rc, name)
    _launch_counter += _LAUNCHES.get(name, 1)
