"""One launch each of the non-GEMM kernels of the hot path at the cfg2 bench shape (64 clips, N=474), between
cudaProfilerStart/Stop, for

    ncu --set full --clock-control none --import-source on --profile-from-start off \
        -o gpurun_out/r2_kernels python tests/ncu_kernels.py

Profiling infrastructure only.  `python tests/ncu_kernels.py time` prints CUDA-event timings instead (no profiler).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from passt_b200 import _lib as L  # noqa: E402
from passt_b200.passt import get_model  # noqa: E402
from passt_b200.preprocess import AugmentMelSTFT  # noqa: E402
from passt_b200 import engine  # noqa: E402

dev = torch.device("cuda:0")
B, H, Dm = int(os.environ.get("NCU_BATCH", "64")), 12, 768
torch.manual_seed(0)
mel = AugmentMelSTFT(freqm=48, timem=192, fmin_aug_range=10, fmax_aug_range=2000).to(dev).train()
net = get_model(arch="passt_s_swa_p16_128_ap476", pretrained=False, s_patchout_t=40, s_patchout_f=4).to(dev).train()
wave = 0.1 * torch.randn(B, 320000, device=dev)
spec = mel(wave).unsqueeze(1)
plan = engine.draw_step_plan(net, spec, True)
ntok = plan.ntok
M = B * ntok
st = L.stream_ptr()
f32 = dict(device=dev, dtype=torch.float32)
b16 = dict(device=dev, dtype=torch.bfloat16)

A0 = torch.empty(M, 256, **b16)
x = torch.randn(M, Dm, **f32)
delta = torch.randn(M, Dm, device=dev).bfloat16()
x_out = torch.empty(M, Dm, **f32)
h = torch.empty(M, Dm, **b16)
mean = torch.empty(M, **f32)
rstd = torch.empty(M, **f32)
gamma = torch.ones(Dm, **f32)
beta = torch.zeros(Dm, **f32)
g = torch.randn(M, Dm, **f32)
gb = torch.empty(M, Dm, **b16)
dgam = torch.zeros(Dm, **f32)
dbet = torch.zeros(Dm, **f32)
csum = torch.zeros(Dm, **f32)
qkv = torch.randn(B, ntok, 3 * Dm, device=dev).bfloat16()
att = torch.empty(B, ntok, Dm, **b16)
npad = ((ntok + 127) // 128) * 128
lse = torch.empty(B, H, npad, **f32)
dO = torch.randn(B, ntok, Dm, device=dev).bfloat16()
dqkv = torch.empty_like(qkv)
dbias = torch.zeros(3 * Dm, **f32)
ws = torch.empty(L.load().passt_attn_bwd_workspace_bytes(B, ntok, H), dtype=torch.uint8, device=dev)
scale = 64 ** -0.5


def k_mel():
    mel(wave)


def k_im2col():
    L.call("passt_im2col", L.ptr(spec), L.ptr(A0), L.ptr(plan.patch_f), L.ptr(plan.patch_t), B, ntok, 128, 1000, 10, 10,
           None, None, st)


def k_ln_fwd():
    L.call("passt_ln_fwd", L.ptr(x), L.ptr(delta), L.ptr(x_out), L.ptr(h), L.ptr(mean), L.ptr(rstd), L.ptr(gamma),
           L.ptr(beta), M, Dm, 1e-6, st)


def k_ln_bwd():
    L.call("passt_ln_bwd", L.ptr(delta), L.ptr(x_out), L.ptr(mean), L.ptr(rstd), L.ptr(gamma), L.ptr(g), L.ptr(g),
           L.ptr(gb), L.ptr(dgam), L.ptr(dbet), L.ptr(csum), M, Dm, st)


wpe = (0.05 * torch.randn(Dm, 256, device=dev)).bfloat16()
tab = torch.randn(ntok, Dm, **f32)
spec32 = spec.contiguous()


def k_patch_embed():
    L.call("passt_patch_embed", L.ptr(spec32), L.ptr(wpe), L.ptr(tab), L.ptr(x_out), L.ptr(plan.patch_f),
           L.ptr(plan.patch_t), B, ntok, 128, 1000, 10, 10, None, None, st)


def k_pe_gemm():
    L.call("passt_gemm_bf16", L.ptr(A0), L.ptr(wpe), L.ptr(x_out), None, None, L.ptr(tab), M, Dm, 256, 256, 256, Dm,
           2, ntok, Dm, 1, 0, st)


def k_attn_fwd():
    L.call("passt_attn_fwd", L.ptr(qkv), L.ptr(att), L.ptr(lse), B, ntok, H, scale, st)


def k_attn_bwd():
    L.call("passt_attn_bwd", L.ptr(qkv), L.ptr(att), L.ptr(dO), L.ptr(lse), L.ptr(dqkv), L.ptr(dbias), L.ptr(ws), B,
           ntok, H, scale, st)


KERNELS = [("mel", k_mel, B * 1.792e6, "B"), ("im2col", k_im2col, B * ((ntok - 2) * 256 * (4 + 2)), "B"),
           ("ln_fwd", k_ln_fwd, M * Dm * 12.0, "B"), ("ln_bwd", k_ln_bwd, M * Dm * 16.0, "B"),
           ("attn_fwd", k_attn_fwd, 4.0 * B * H * ntok * ntok * 64, "F"),
           ("attn_bwd", k_attn_bwd, 10.0 * B * H * ntok * ntok * 64, "F"),
           ("patch_embed", k_patch_embed, B * (ntok - 2) * 1024.0 + M * Dm * 4.0, "B"),
           ("pe_gemm", k_pe_gemm, M * 256 * 2.0 + M * Dm * 4.0, "B")]
if os.environ.get("NCU_ONLY"):
    KERNELS = [k for k in KERNELS if k[0] in os.environ["NCU_ONLY"].split(",")]

for _, fn, _, _ in KERNELS:          # warm-up (cudaFuncSetAttribute, table builds)
    fn()
    fn()
torch.cuda.synchronize()

if sys.argv[1:] == ["time"]:
    for name, fn, work, kind in KERNELS:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        rate = work / (ms * 1e-3)
        print(f"{name}: {ms * 1e3:.1f} us  " + (f"{rate / 1e9:.0f} GB/s algorithmic" if kind == "B" else f"{rate / 1e12:.0f} TFLOP/s"))
    sys.exit(0)

torch.cuda.cudart().cudaProfilerStart()
for _, fn, _, _ in KERNELS:
    fn()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok", "ntok", ntok)
