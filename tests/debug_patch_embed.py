"""Bring-up of passt_patch_embed alone (no engine): small shapes, sync after the launch, compare with torch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from passt_b200 import _lib as L

dev = torch.device("cuda:0")
lib = L.load()
lib.passt_set_pdl(int(os.environ.get("DBG_PDL", "0")))
torch.manual_seed(0)


def run(B, Fg, Tg, T, mix):
    ntok = Fg * Tg + 2
    mel = torch.randn(B, 128, T, device=dev)
    pf = torch.arange(Fg).repeat_interleave(Tg).to(torch.int32).to(dev)
    pt = torch.arange(Tg).repeat(Fg).to(torch.int32).to(dev)
    W = (torch.randn(768, 256, device=dev) * 0.05)
    Wb = W.bfloat16()
    tab = torch.randn(ntok, 768, device=dev)
    out = torch.full((B * ntok, 768), float("nan"), device=dev)
    perm = lam = None
    if mix:
        perm = torch.randperm(B).to(torch.int32).to(dev)
        lam = (torch.rand(B) * 0.5 + 0.5).to(dev)
    L.call("passt_patch_embed", L.ptr(mel), L.ptr(Wb), L.ptr(tab), L.ptr(out), L.ptr(pf), L.ptr(pt), B, ntok, 128, T, 10, 10,
           L.ptr(perm), L.ptr(lam), L.stream_ptr())
    torch.cuda.synchronize()
    # reference
    x = mel
    if mix:
        l = lam.view(B, 1, 1)
        x = mel * l + mel[perm.long()] * (1 - l)
    patches = torch.stack([x[:, f * 10: f * 10 + 16, t * 10: t * 10 + 16].reshape(B, 256)
                           for f, t in zip(pf.tolist(), pt.tolist())], 1)          # [B, ntok-2, 256]
    emb = patches.bfloat16().float() @ Wb.float().t()
    ref = torch.cat([torch.zeros(B, 2, 768, device=dev), emb], 1) + tab.unsqueeze(0)
    err = (out.view(B, ntok, 768) - ref).abs().max().item() / ref.abs().max().item()
    print(f"B={B} grid={Fg}x{Tg} ntok={ntok} T={T} mix={mix}: relerr {err:.2e}", flush=True)


for args in [(1, 4, 10, 160, False), (1, 12, 10, 200, False), (2, 12, 20, 256, False), (3, 8, 59, 1000, True), (5, 12, 99, 1000, False)]:
    run(*args)
print("done")
