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


def grad_metrics(a, b):
    """Error metrics of candidate `a` against reference `b` (same shape):
      relmax  max|a-b| / max|b|                     (the whole-tensor norm north_star's tolerances are stated in)
      rel_l2  ||a-b||_2 / ||b||_2
      cos     cosine similarity of the flattened tensors
      elrel   max over elements with |b| > 1e-3 max|b| of |a-b| / |b|   (element-relative; bf16 noise shows here)
      elrel50 median of the same ratio
    A tensor whose small entries are wrong passes `relmax` but not `rel_l2` / `cos`."""
    a = a.detach().double().cpu().flatten()
    b = b.detach().double().cpu().flatten()
    d = (a - b).abs()
    bmax = b.abs().max().item() + 1e-300
    sel = b.abs() > 1e-3 * bmax
    ratio = d[sel] / b[sel].abs() if bool(sel.any()) else torch.zeros(1, dtype=torch.float64)
    return dict(relmax=d.max().item() / bmax, rel_l2=(a - b).norm().item() / (b.norm().item() + 1e-300),
                cos=float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-300)), elrel=ratio.max().item(),
                elrel50=ratio.median().item(), refmax=bmax)


def golden_sample_index(n: int, n_samples: int = 4096):
    return torch.arange(n)[:: max(1, n // n_samples)][:n_samples]


def golden_projections(name: str, g, n_proj: int = 4):
    """Same random projections as tests/golden/make_golden_grads.py (seeded by crc32 of the parameter name)."""
    import zlib
    out = []
    flat = g.detach().double().cpu().flatten()
    for j in range(n_proj):
        gen = torch.Generator().manual_seed((zlib.crc32(name.encode()) + j) & 0x7FFFFFFF)
        r = torch.randn(flat.numel(), generator=gen, dtype=torch.float64)
        out.append(float(flat @ r))
    return torch.tensor(out, dtype=torch.float64)
