import contextlib
import io
import os
import sys
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        yield


def relerr(a, b):
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-20)).item()


def build_net(cfg, params, device, arch="passt_s_swa_p16_128_ap476", cut_depth=0):
    """Candidate net with oracle-format params loaded (tests only)."""
    from passt_b200 import passt as P
    with quiet():
        net = P.get_model(arch=arch, pretrained=False, n_classes=cfg.n_classes, fstride=cfg.fstride,
                          tstride=cfg.tstride, input_fdim=cfg.input_fdim, input_tdim=cfg.input_tdim,
                          u_patchout=cfg.u_patchout, s_patchout_t=cfg.s_patchout_t, s_patchout_f=cfg.s_patchout_f)
    sd = {k: v for k, v in params.items()}
    net.load_state_dict(sd, strict=True)
    if cut_depth:
        net = P.lighten_model(net, cut_depth=cut_depth)
    return net.to(device)


def oracle_cfg_for_depth(depth, **kw):
    from oracle import passt_oracle as O
    return O.NetCfg(depth=depth, **kw)


def depth2_params(params12, last=11):
    """Oracle-format parameters of the 2-block network lighten_model(cut_depth=10) leaves: block 0 and block `last`
    (models/passt.py:932-954), renumbered 0, 1."""
    out = {}
    for k, v in params12.items():
        if k.startswith("blocks."):
            i = int(k.split(".")[1])
            if i == 0:
                out[k] = v
            elif i == last:
                out[k.replace(f"blocks.{last}.", "blocks.1.")] = v
        else:
            out[k] = v
    return out
