"""One train step of the bench workload between cudaProfilerStart/Stop (for `ncu --profile-from-start off`).
Test/profiling infrastructure only."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from passt_b200.passt import get_model  # noqa: E402
from passt_b200.preprocess import AugmentMelSTFT  # noqa: E402
from passt_b200 import loss as PL  # noqa: E402


def main():
    B = int(os.environ.get("PROFILE_BATCH", "64"))
    steps = int(os.environ.get("PROFILE_STEPS", "1"))
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    mel = AugmentMelSTFT(freqm=48, timem=192, fmin_aug_range=10, fmax_aug_range=2000).to(dev).train()
    net = get_model(arch="passt_s_swa_p16_128_ap476", pretrained=False, s_patchout_t=40, s_patchout_f=4).to(dev).train()
    from passt_b200.optim import FusedAdamW
    opt = FusedAdamW([p for n, p in net.named_parameters() if not n.startswith("head_dist")], lr=2e-5,
                     weight_decay=1e-4).attach(net)          # same optimizer as bench.py
    wave = 0.1 * torch.randn(B, 320000, device=dev)
    y = (torch.rand(B, 527, device=dev) < 0.005).float()

    def step():
        with torch.no_grad():
            spec = mel(wave).unsqueeze(1)
        logits, _ = net(spec)
        loss = PL.bce_with_logits(logits, y)            # fused loss kernel, as bench.py
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if os.environ.get("PROFILE_CPU"):
        import time
        t0 = time.perf_counter()
        for _ in range(5):
            step()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"cpu enqueue per step {(t1 - t0) / 5 * 1e3:.2f} ms; wall per step incl. drain {(t2 - t0) / 5 * 1e3:.2f} ms")
    torch.cuda.cudart().cudaProfilerStart()
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
