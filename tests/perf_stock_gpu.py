"""Stock PyTorch-CUDA train step of the reference algorithm on the same B200 (denominator of north_star's ">= 3x").

The reference modules cannot travel to the GPU box (/root/reference is absent there), so this times the oracle
port — bit-exact with the reference on CPU (tests/test_oracle_vs_reference.py) and made of the very same torch ops
the reference calls (torch.stft-equivalent rfft, F.conv2d, F.layer_norm, F.linear, softmax, F.gelu) — on CUDA under
autocast, eager and torch.compile'd, protocol of model_speed_test (ex_audioset.py:364-426) extended to start from
waveforms.  Test infrastructure only (lives under tests/, prints JSON lines to stdout).

    python tests/perf_stock_gpu.py [--steps 20] [--batch 64]
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import passt_oracle as O  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--compile", type=int, default=1)
    args = ap.parse_args()
    dev = "cuda"
    mcfg = O.MelCfg()
    ncfg = O.NetCfg(s_patchout_t=40, s_patchout_f=4)
    params = {k: v.to(dev).requires_grad_(not k.startswith("head_dist")) for k, v in O.synth_params(ncfg, 0).items()}
    opt = torch.optim.AdamW([p for p in params.values() if p.requires_grad], lr=2e-5, weight_decay=1e-4, fused=True)
    B = args.batch
    torch.manual_seed(0)
    wave = 0.1 * torch.randn(B, 320000, device=dev)
    y = (torch.rand(B, 527, device=dev) < 0.005).float()

    def net_fn(spec, t_keep, f_keep):
        d = O.StepDraws(t_keep=t_keep, f_keep=f_keep)
        return O.passt_forward(params, spec, ncfg, d)[0]

    for name, dtype, fn in [("eager_bf16", torch.bfloat16, net_fn), ("eager_fp16", torch.float16, net_fn),
                            ("compiled_bf16", torch.bfloat16, torch.compile(net_fn) if args.compile else None)]:
        if fn is None:
            continue
        scaler = torch.amp.GradScaler("cuda", enabled=(dtype == torch.float16))

        def step():
            d = O.draw_mel(mcfg, True, B, device=dev)
            with torch.no_grad():
                spec = O.mel_frontend(wave, mcfg, d, True).unsqueeze(1)
            dp = O.draw_patchout(ncfg, 12, 99, True)
            with torch.autocast("cuda", dtype=dtype):
                logits = fn(spec, dp.t_keep.to(dev), dp.f_keep.to(dev))
            loss = F.binary_cross_entropy_with_logits(logits.float(), y)
            opt.zero_grad(set_to_none=True)
            scaler.scale(loss).backward()
            scaler.step(opt)
            scaler.update()

        try:
            for _ in range(5):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            for _ in range(args.steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            print(json.dumps({"stock": name, "batch": B, "ms_per_step": ms, "clips_per_s": B / ms * 1e3,
                              "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}), flush=True)
        except Exception as e:  # noqa
            print(json.dumps({"stock": name, "error": str(e)[:300]}), flush=True)


if __name__ == "__main__":
    main()
