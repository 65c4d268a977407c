"""How far does the REFERENCE ALGORITHM ITSELF move when torch runs it in bf16?  (tests/golden/bf16_floor.json)

Runs the oracle port (bit-exact with the reference on CPU) on the full-depth gradient fixture case twice -- fp32, and
under ``torch.autocast("cpu", dtype=torch.bfloat16)`` (bf16 GEMM operands, fp32 accumulation: the arithmetic of the
reference's mixed-precision training) -- and records the per-parameter gradient error of the second against the first.
This is the noise floor any bf16 tensor-core implementation of the 12-block backward sits on; the GPU parity tests hold
the CUDA path to it (tests/test_gpu_fulldepth.py).  CPU only:  python tests/golden/make_bf16_floor.py
"""
import json
import os
import statistics
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import passt_oracle as O  # noqa: E402
from util import grad_metrics  # noqa: E402


def main():
    G = torch.load(os.path.join(HERE, "passt_golden_grads.pt"))
    cfg = O.NetCfg(**G["net_kw"])

    def run(autocast):
        p = {k: v.clone().requires_grad_(True) for k, v in O.synth_params(cfg, seed=G["param_seed"]).items()}
        torch.manual_seed(G["input_seed"])
        x = torch.randn(*G["x_shape"])
        torch.manual_seed(G["rng_seed"])
        d = O.draw_patchout(cfg, 12, 99, True)
        if autocast:
            with torch.autocast("cpu", dtype=torch.bfloat16):
                lg, _ = O.passt_forward(p, x, cfg, d)
        else:
            lg, _ = O.passt_forward(p, x, cfg, d)
        torch.manual_seed(G["grad_weight_seed"])
        w = torch.randn(lg.shape)
        (lg.float() * w).sum().backward()
        return lg.detach().float(), {k: v.grad for k, v in p.items() if v.grad is not None}

    l0, g0 = run(False)
    l1, g1 = run(True)
    per = {k: grad_metrics(g1[k], g0[k]) for k in g0}
    rel = [m["relmax"] for m in per.values()]
    out = {"what": "oracle port under torch.autocast('cpu', bfloat16) vs the same in fp32; case of passt_golden_grads.pt",
           "torch": torch.__version__, "logits_relmax": float((l1 - l0).abs().max() / l0.abs().max()),
           "worst_relmax": max(rel), "median_relmax": statistics.median(rel), "n_over_1e-2": sum(r > 1e-2 for r in rel),
           "n_tensors": len(rel), "worst_rel_l2": max(m["rel_l2"] for m in per.values()),
           "min_cos": min(m["cos"] for m in per.values()),
           "per_parameter_relmax": {k: m["relmax"] for k, m in per.items()}}
    with open(os.path.join(HERE, "bf16_floor.json"), "w") as f:
        json.dump(out, f, indent=1)
    print({k: v for k, v in out.items() if k != "per_parameter_relmax"})


if __name__ == "__main__":
    main()
