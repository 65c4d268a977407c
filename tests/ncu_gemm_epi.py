"""One launch each of the epilogue-heavy GEMM modes at the bench shape, for `ncu --set full --import-source on`.

usage (GPU box): ncu --set full --import-source on --clock-control none -k regex:gemm2_kernel -o gpurun_out/gemm_epi python tests/ncu_gemm_epi.py
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from passt_b200 import _lib as L

dev = torch.device("cuda:0")
M, N, K = 30336, 3072, 768
torch.manual_seed(0)
A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
W = (torch.randn(N, K, device=dev) * 0.05).bfloat16()          # [N, K] K-major
Wkn = W.t().contiguous()                                         # [K, N]
bias = torch.randn(N, device=dev)
C = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
C2 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
aux = torch.randn(M, N, device=dev).bfloat16()
cs = torch.zeros(N, device=dev)
st = L.stream_ptr()


def gemm(Bm, ldb, mode, C, C2=None, bias=None, aux=None):
    L.call("passt_gemm_bf16", L.ptr(A), L.ptr(Bm), L.ptr(C), L.ptr(C2), L.ptr(bias), L.ptr(aux), M, N, K, K, ldb, N,
           mode, 0, N, 1, 0, st)


which = sys.argv[1:] or ["0", "1", "3"]
if which == ["time"]:
    # plain event timing (no profiler): us per launch and TFLOP/s of each mode at this shape
    def timeit(fn, n=30):
        for _ in range(5):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    for name, fn in (("mode0", lambda: gemm(W, K, 0, C, bias=bias)), ("mode1", lambda: gemm(W, K, 1, C, C2=C2, bias=bias)),
                     ("mode0_bkn", lambda: gemm(Wkn, N, 16, C, bias=bias)),
                     ("mode3_bkn", lambda: gemm(Wkn, N, 3 | 16, C, bias=cs, aux=aux))):
        us = timeit(fn)
        print(f"{name}: {us:.1f} us  {2.0 * M * N * K / us / 1e6:.0f} TFLOP/s")
    # mode-3 numerics against torch
    cs.zero_()
    gemm(Wkn, N, 3 | 16, C, bias=cs, aux=aux)
    torch.cuda.synchronize()
    ref = (A.float() @ Wkn.float()) * aux.float()
    print("mode3 relerr", ((C.float() - ref).norm() / ref.norm()).item(), "colsum relerr",
          ((cs - ref.sum(0)).norm() / ref.sum(0).norm()).item())
    sys.exit(0)
if "0" in which:
    gemm(W, K, 0, C, bias=bias)
if "1" in which:
    gemm(W, K, 1, C, C2=C2, bias=bias)
if "3" in which:
    gemm(Wkn, N, 3 | 16, C, bias=cs, aux=aux)
torch.cuda.synchronize()
print("ok")
