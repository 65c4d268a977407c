"""GPU bring-up harness (not a pytest file): python tests/bringup_gpu.py <group> ; appends JSON lines to
gpurun_out/bringup.jsonl.  Each group runs in its own process so that a trapping kernel cannot poison the rest."""
import json
import os
import sys
import time
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from passt_b200 import _lib as L  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
dev = "cuda"


def log(**kw):
    kw["t"] = time.time()
    line = json.dumps(kw, default=str)
    print(line, flush=True)
    with open(os.path.join(OUT, "bringup.jsonl"), "a") as f:
        f.write(line + "\n")


def relerr(a, b):
    a = a.float(); b = b.float()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def gemm(A, B, C, C2=None, bias=None, aux=None, M=0, N=0, K=0, lda=0, ldb=0, ldc=0, mode=0, period=0, ld_aux=0,
         splits=1, max_ctas=0):
    L.call("passt_gemm_bf16", L.ptr(A), L.ptr(B), L.ptr(C), L.ptr(C2), L.ptr(bias), L.ptr(aux), M, N, K, lda, ldb,
           ldc, mode, period, ld_aux, splits, max_ctas, L.stream_ptr())


def timeit(fn, iters=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def g_gemm_tn():
    torch.manual_seed(0)
    for (M, N, K) in [(128, 256, 64), (300, 256, 128), (1000, 768, 768), (30336, 2304, 768), (30336, 768, 3072)]:
        A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
        B = (torch.randn(N, K, device=dev) * 0.5).bfloat16()
        bias = torch.randn(N, device=dev)
        C = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16)
        ref = A.float() @ B.float().t() + bias
        gemm(A, B, C, bias=bias, M=M, N=N, K=K, lda=K, ldb=K, ldc=N, mode=0)
        torch.cuda.synchronize()
        err = relerr(C, ref)
        rec = dict(test="gemm_tn_bias", M=M, N=N, K=K, relerr=err, ok=bool(err < 2e-2), nan=int(torch.isnan(C.float()).sum()))
        if M >= 30000:
            ms = timeit(lambda: gemm(A, B, C, bias=bias, M=M, N=N, K=K, lda=K, ldb=K, ldc=N, mode=0))
            rec["ms"] = ms; rec["tflops"] = 2.0 * M * N * K / ms / 1e9
        log(**rec)
        if err > 2e-2 and M <= 300:
            # diagnostics: dump a corner and probe descriptor alternatives
            log(test="gemm_tn_diag", got=C[:4, :8].float().tolist(), want=ref[:4, :8].tolist())
            import ctypes
            for alt in [(0, 1024, 32), (1, 1024, 32), (1024, 1024, 32), (16, 1024, 2), (16, 64, 32), (16, 1024, 64)]:
                arr = (ctypes.c_uint * 6)(alt[0], alt[1], alt[2], alt[0], alt[1], alt[2])
                L.load().passt_gemm_debug_desc(1, arr)
                C.fill_(float("nan"))
                try:
                    gemm(A, B, C, bias=bias, M=M, N=N, K=K, lda=K, ldb=K, ldc=N, mode=0)
                    torch.cuda.synchronize()
                    log(test="gemm_tn_alt", alt=alt, relerr=relerr(C, ref))
                except Exception as e:  # noqa
                    log(test="gemm_tn_alt", alt=alt, error=str(e))
                    break
            L.load().passt_gemm_debug_desc(0, None)


def g_gemm_modes():
    torch.manual_seed(1)
    M, N, K = 1000, 768, 256
    A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    B = (torch.randn(N, K, device=dev) * 0.5).bfloat16()
    bias = torch.randn(N, device=dev)
    acc = A.float() @ B.float().t()
    # mode 1: bias + gelu dual
    N1 = 3072
    B1 = (torch.randn(N1, K, device=dev) * 0.2).bfloat16()
    b1 = torch.randn(N1, device=dev)
    C = torch.empty(M, N1, device=dev, dtype=torch.bfloat16); C2 = torch.empty_like(C)
    gemm(A, B1, C, C2=C2, bias=b1, M=M, N=N1, K=K, lda=K, ldb=K, ldc=N1, mode=1)
    torch.cuda.synchronize()
    pre = (A.float() @ B1.float().t() + b1).requires_grad_(True)
    act_ref = torch.nn.functional.gelu(pre)
    act_ref.sum().backward()
    log(test="gemm_mode1", dgelu=relerr(C, pre.grad), post=relerr(C2, act_ref))
    # mode 2: row table fp32
    period = 250
    tab = torch.randn(period, N, device=dev)
    Cf = torch.empty(M, N, device=dev)
    gemm(A, B, Cf, aux=tab, M=M, N=N, K=K, lda=K, ldb=K, ldc=N, mode=2, period=period, ld_aux=N)
    torch.cuda.synchronize()
    ref = acc + tab.repeat(M // period, 1)
    log(test="gemm_mode2", relerr=relerr(Cf, ref))
    # mode 3: gelu-grad
    prev = (torch.randn(M, N, device=dev)).bfloat16()
    Cb = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    csum = torch.zeros(N, device=dev)
    gemm(A, B, Cb, aux=prev, bias=csum, M=M, N=N, K=K, lda=K, ldb=K, ldc=N, mode=3, ld_aux=N)
    torch.cuda.synchronize()
    log(test="gemm_mode3", relerr=relerr(Cb, acc * prev.float()), colsum=relerr(csum, (acc * prev.float()).sum(0)))


def g_gemm_wgrad():
    import ctypes
    torch.manual_seed(2)
    for (Kt, M, N, splits) in [(64, 128, 256, 1), (200, 256, 256, 1), (1000, 768, 768, 4), (30336, 2304, 768, 16),
                               (30336, 768, 3072, 8)]:
        A = (torch.randn(Kt, M, device=dev) * 0.5).bfloat16()
        B = (torch.randn(Kt, N, device=dev) * 0.5).bfloat16()
        C = torch.zeros(M, N, device=dev)
        ref = A.float().t() @ B.float()
        gemm(A, B, C, M=M, N=N, K=Kt, lda=M, ldb=N, ldc=N, mode=4, splits=splits)
        torch.cuda.synchronize()
        err = relerr(C, ref)
        rec = dict(test="gemm_wgrad", Kt=Kt, M=M, N=N, splits=splits, relerr=err, ok=bool(err < 2e-2))
        if Kt >= 30000:
            ms = timeit(lambda: gemm(A, B, C, M=M, N=N, K=Kt, lda=M, ldb=N, ldc=N, mode=4, splits=splits))
            rec["ms"] = ms; rec["tflops"] = 2.0 * M * N * Kt / ms / 1e9
        log(**rec)
        if err > 2e-2 and Kt <= 200:
            log(test="gemm_wgrad_diag", got=C[:4, :8].tolist(), want=ref[:4, :8].tolist())
            for alt in [(1024, 8192, 2048), (8192, 1024, 1024), (8192, 1024, 256), (8192, 128, 2048),
                        (128, 8192, 2048), (8192, 1024, 32)]:
                arr = (ctypes.c_uint * 6)(alt[0], alt[1], alt[2], alt[0], alt[1], alt[2])
                L.load().passt_gemm_debug_desc(1, arr)
                C.zero_()
                try:
                    gemm(A, B, C, M=M, N=N, K=Kt, lda=M, ldb=N, ldc=N, mode=4, splits=splits)
                    torch.cuda.synchronize()
                    log(test="gemm_wgrad_alt", alt=alt, relerr=relerr(C, ref))
                except Exception as e:  # noqa
                    log(test="gemm_wgrad_alt", alt=alt, error=str(e))
                    break
            L.load().passt_gemm_debug_desc(0, None)


def g_mel():
    from oracle import passt_oracle as O
    cfg = O.MelCfg()
    torch.manual_seed(0)
    B, Lw = 2, 320000
    wave = 0.1 * torch.randn(B, Lw)
    ws = torch.empty(L.load().passt_mel_workspace_bytes(), dtype=torch.uint8, device=dev)
    L.call("passt_mel_init", L.ptr(ws), cfg.win_length, L.stream_ptr())
    wd = wave.to(dev)
    for training in (False, True):
        torch.manual_seed(7)
        d = O.draw_mel(cfg, training, B)
        ref = O.mel_frontend(wave, cfg, d, training)
        L.call("passt_mel_set_band", L.ptr(ws), float(d.fmin), float(d.fmax), cfg.sr, L.stream_ptr())
        out = torch.empty(B, 128, 1000, device=dev)
        rnd = d.mask_rnd.to(dev).contiguous() if training else None
        L.call("passt_mel_forward", L.ptr(ws), L.ptr(wd), L.ptr(out), B, Lw, cfg.hopsize, L.ptr(rnd),
               cfg.freqm if training else 0, cfg.timem if training else 0, L.stream_ptr())
        torch.cuda.synchronize()
        diff = (out.cpu() - ref).abs()
        log(test="mel", training=training, maxabs=diff.max().item(), mean=diff.mean().item(),
            fmin=d.fmin, fmax=d.fmax, n_bad=int((diff > 1e-3).sum()))
    Bb = 64
    wv = 0.1 * torch.randn(Bb, Lw, device=dev)
    out = torch.empty(Bb, 128, 1000, device=dev)
    ms = timeit(lambda: L.call("passt_mel_forward", L.ptr(ws), L.ptr(wv), L.ptr(out), Bb, Lw, 320, None, 0, 0,
                               L.stream_ptr()))
    log(test="mel_time", B=Bb, ms=ms, clips_per_s=Bb / ms * 1e3, gbps=Bb * 1.792e6 / ms / 1e6)


def g_rowops():
    torch.manual_seed(3)
    M, Dm = 1000, 768
    x = torch.randn(M, Dm, device=dev)
    delta = (0.3 * torch.randn(M, Dm, device=dev)).bfloat16()
    g = 1 + 0.1 * torch.randn(Dm, device=dev); b = 0.1 * torch.randn(Dm, device=dev)
    xo = torch.empty_like(x); h = torch.empty(M, Dm, device=dev, dtype=torch.bfloat16)
    mean = torch.empty(M, device=dev); rstd = torch.empty(M, device=dev)
    L.call("passt_ln_fwd", L.ptr(x), L.ptr(delta), L.ptr(xo), L.ptr(h), L.ptr(mean), L.ptr(rstd), L.ptr(g), L.ptr(b),
           M, Dm, 1e-6, L.stream_ptr())
    xr = (x + delta.float()).requires_grad_(True)
    gr = g.clone().requires_grad_(True); br = b.clone().requires_grad_(True)
    hr = torch.nn.functional.layer_norm(xr, (Dm,), gr, br, 1e-6)
    log(test="ln_fwd", x=relerr(xo, xr), h=relerr(h, hr))
    dh = (torch.randn(M, Dm, device=dev)).bfloat16()
    gin = torch.randn(M, Dm, device=dev)
    hr.backward(dh.float())
    gout = torch.empty_like(x); goutb = torch.empty(M, Dm, device=dev, dtype=torch.bfloat16)
    dg = torch.zeros(Dm, device=dev); db = torch.zeros(Dm, device=dev); cs = torch.zeros(Dm, device=dev)
    L.call("passt_ln_bwd", L.ptr(dh), L.ptr(xo), L.ptr(mean), L.ptr(rstd), L.ptr(g), L.ptr(gin), L.ptr(gout),
           L.ptr(goutb), L.ptr(dg), L.ptr(db), L.ptr(cs), M, Dm, L.stream_ptr())
    torch.cuda.synchronize()
    ref_g = gin + xr.grad
    log(test="ln_bwd", gout=relerr(gout, ref_g), goutb=relerr(goutb, ref_g), dgamma=relerr(dg, gr.grad),
        dbeta=relerr(db, br.grad), colsum=relerr(cs, goutb.float().sum(0)))
    # colsum
    Mx, Nx = 3001, 2304
    inp = torch.randn(Mx, Nx, device=dev).bfloat16()
    out = torch.zeros(Nx, device=dev)
    L.call("passt_colsum_bf16", L.ptr(inp), L.ptr(out), Mx, Nx, Nx, L.stream_ptr())
    log(test="colsum", relerr=relerr(out, inp.float().sum(0)))
    # cast_transpose
    W = torch.randn(300, 500, device=dev)
    o1 = torch.empty(300, 500, device=dev, dtype=torch.bfloat16); o2 = torch.empty(500, 300, device=dev, dtype=torch.bfloat16)
    L.call("passt_cast_transpose", L.ptr(W), L.ptr(o1), L.ptr(o2), 300, 500, L.stream_ptr())
    log(test="cast_transpose", a=relerr(o1, W.bfloat16()), b=relerr(o2, W.t().bfloat16()))
    # im2col + token table
    B, Fm, Tm = 2, 128, 1000
    mel = torch.randn(B, Fm, Tm, device=dev)
    pf = torch.randint(0, 12, (30,), device=dev, dtype=torch.int32)
    pt = torch.randint(0, 99, (30,), device=dev, dtype=torch.int32)
    ntok = 32
    A = torch.empty(B * ntok, 256, device=dev, dtype=torch.bfloat16)
    L.call("passt_im2col", L.ptr(mel), L.ptr(A), L.ptr(pf), L.ptr(pt), B, ntok, Fm, Tm, 10, 10, None, None,
           L.stream_ptr())
    ref = torch.zeros(B, ntok, 256, device=dev)
    for n in range(2, ntok):
        f0, t0 = int(pf[n - 2]) * 10, int(pt[n - 2]) * 10
        ref[:, n] = mel[:, f0:f0 + 16, t0:t0 + 16].reshape(B, 256)
    log(test="im2col", relerr=relerr(A, ref.reshape(B * ntok, 256).bfloat16()))
    cls = torch.randn(Dm, device=dev); dist = torch.randn(Dm, device=dev); npos = torch.randn(2, Dm, device=dev)
    cb = torch.randn(Dm, device=dev); tp = torch.randn(Dm, 99, device=dev); fp = torch.randn(Dm, 12, device=dev)
    tab = torch.empty(ntok, Dm, device=dev)
    L.call("passt_token_table", L.ptr(tab), L.ptr(cls), L.ptr(dist), L.ptr(npos), L.ptr(cb), L.ptr(tp), L.ptr(fp),
           L.ptr(pf), L.ptr(pt), ntok, 12, 99, 0, None, L.stream_ptr())
    reft = torch.empty(ntok, Dm, device=dev)
    reft[0] = cls + npos[0]; reft[1] = dist + npos[1]
    for n in range(2, ntok):
        reft[n] = cb + tp[:, int(pt[n - 2])] + fp[:, int(pf[n - 2])]
    log(test="token_table", relerr=relerr(tab, reft))
    g0 = torch.randn(B, ntok, Dm, device=dev)
    outs = [torch.zeros(Dm, device=dev), torch.zeros(Dm, device=dev), torch.zeros(2, Dm, device=dev),
            torch.zeros(Dm, device=dev), torch.zeros(Dm, 99, device=dev), torch.zeros(Dm, 12, device=dev)]
    L.call("passt_token_table_bwd", L.ptr(g0), *[L.ptr(o) for o in outs], L.ptr(pf), L.ptr(pt), B, ntok, 12, 99, 0,
           None, L.stream_ptr())
    s = g0.sum(0)
    rt = torch.zeros(Dm, 99, device=dev); rf = torch.zeros(Dm, 12, device=dev)
    for n in range(2, ntok):
        rt[:, int(pt[n - 2])] += s[n]; rf[:, int(pf[n - 2])] += s[n]
    log(test="token_table_bwd", cls=relerr(outs[0], s[0]), dist=relerr(outs[1], s[1]),
        npos=relerr(outs[2], s[:2]), cb=relerr(outs[3], s[2:].sum(0)), time=relerr(outs[4], rt),
        freq=relerr(outs[5], rf))
    # head
    Bh, nt, Cc = 4, 20, 527
    xh = torch.randn(Bh, nt, Dm, device=dev)
    dl = (0.2 * torch.randn(Bh, nt, Dm, device=dev)).bfloat16()
    ng = 1 + 0.1 * torch.randn(Dm, device=dev); nb = 0.1 * torch.randn(Dm, device=dev)
    hg = 1 + 0.1 * torch.randn(Dm, device=dev); hb = 0.1 * torch.randn(Dm, device=dev)
    Wh = 0.05 * torch.randn(Cc, Dm, device=dev); bh = 0.1 * torch.randn(Cc, device=dev)
    logits = torch.empty(Bh, Cc, device=dev); feats = torch.empty(Bh, Dm, device=dev); fl = torch.empty(Bh, Dm, device=dev)
    L.call("passt_head_fwd", L.ptr(xh), L.ptr(dl), L.ptr(ng), L.ptr(nb), L.ptr(hg), L.ptr(hb), L.ptr(Wh), L.ptr(bh),
           L.ptr(logits), L.ptr(feats), L.ptr(fl), Bh, nt, Cc, L.stream_ptr())
    leaves = [t.clone().requires_grad_(True) for t in ((xh + dl.float()), ng, nb, hg, hb, Wh, bh)]
    xv, ng_, nb_, hg_, hb_, Wh_, bh_ = leaves
    y = torch.nn.functional.layer_norm(xv, (Dm,), ng_, nb_, 1e-6)
    fr = (y[:, 0] + y[:, 1]) / 2
    lr = torch.nn.functional.linear(torch.nn.functional.layer_norm(fr, (Dm,), hg_, hb_, 1e-5), Wh_, bh_)
    log(test="head_fwd", logits=relerr(logits, lr), feats=relerr(feats, fr))
    dlog = torch.randn(Bh, Cc, device=dev); dfe = torch.randn(Bh, Dm, device=dev)
    (lr * dlog).sum().backward(retain_graph=True); (fr * dfe).sum().backward()
    gout = torch.zeros(Bh, nt, Dm, device=dev); goutb = torch.zeros(Bh, nt, Dm, device=dev, dtype=torch.bfloat16)
    z = lambda *s: torch.zeros(*s, device=dev)
    dng, dnb, dhg, dhb, dW, dbh, cs = z(Dm), z(Dm), z(Dm), z(Dm), z(Cc, Dm), z(Cc), z(Dm)
    L.call("passt_head_bwd", L.ptr(xh), L.ptr(dl), L.ptr(ng), L.ptr(nb), L.ptr(hg), L.ptr(hb), L.ptr(Wh), L.ptr(dlog),
           L.ptr(dfe), L.ptr(fl), L.ptr(gout), L.ptr(goutb), L.ptr(dng), L.ptr(dnb), L.ptr(dhg), L.ptr(dhb), L.ptr(dW),
           L.ptr(dbh), L.ptr(cs), Bh, nt, Cc, L.stream_ptr())
    torch.cuda.synchronize()
    log(test="head_bwd", g=relerr(gout, xv.grad), dng=relerr(dng, ng_.grad), dnb=relerr(dnb, nb_.grad),
        dhg=relerr(dhg, hg_.grad), dhb=relerr(dhb, hb_.grad), dW=relerr(dW, Wh_.grad), dbh=relerr(dbh, bh_.grad),
        cs=relerr(cs, goutb.float().sum((0, 1))))


def g_attn():
    torch.manual_seed(4)
    H, hd = 12, 64
    C = H * hd
    scale = hd ** -0.5
    for (B, N) in [(2, 128), (2, 130), (2, 474), (1, 1190), (64, 474)]:
        qkv = (torch.randn(B, N, 3 * C, device=dev)).bfloat16()
        out = torch.full((B, N, C), float("nan"), device=dev, dtype=torch.bfloat16)
        Npad = ((N + 127) // 128) * 128
        lse = torch.empty(B, H, Npad, device=dev)
        L.call("passt_attn_fwd", L.ptr(qkv), L.ptr(out), L.ptr(lse), B, N, H, scale, L.stream_ptr())
        torch.cuda.synchronize()
        rec = dict(test="attn_fwd", B=B, N=N)
        if B <= 2:
            x = qkv.float().reshape(B, N, 3, H, hd).permute(2, 0, 3, 1, 4)
            q, k, v = [t.clone().requires_grad_(True) for t in (x[0], x[1], x[2])]
            att = (q @ k.transpose(-2, -1)) * scale
            ref_lse = torch.logsumexp(att, -1)
            ref = (att.softmax(-1) @ v).transpose(1, 2).reshape(B, N, C)
            rec.update(o=relerr(out, ref), lse=relerr(lse[:, :, :N] * 0.6931471805599453, ref_lse),
                       nan=int(torch.isnan(out.float()).sum()))
        else:
            ms = timeit(lambda: L.call("passt_attn_fwd", L.ptr(qkv), L.ptr(out), L.ptr(lse), B, N, H, scale, L.stream_ptr()))
            rec.update(ms=ms, tflops=4.0 * B * H * N * N * hd / ms / 1e9)
        log(**rec)
        # backward
        dO = (torch.randn(B, N, C, device=dev)).bfloat16()
        dqkv = torch.full((B, N, 3 * C), float("nan"), device=dev, dtype=torch.bfloat16)
        wsb = L.load().passt_attn_bwd_workspace_bytes(B, N, H)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        dbias = torch.zeros(3 * C, device=dev)
        L.call("passt_attn_bwd", L.ptr(qkv), L.ptr(out), L.ptr(dO), L.ptr(lse), L.ptr(dqkv), L.ptr(dbias), L.ptr(ws), B, N,
               H, scale, L.stream_ptr())
        torch.cuda.synchronize()
        rec = dict(test="attn_bwd", B=B, N=N, dbias=relerr(dbias, dqkv.float().sum((0, 1))))
        if B <= 2:
            ref.backward(dO.float())
            g = dqkv.float().reshape(B, N, 3, H, hd).permute(2, 0, 3, 1, 4)
            rec.update(dq=relerr(g[0], q.grad), dk=relerr(g[1], k.grad), dv=relerr(g[2], v.grad),
                       nan=int(torch.isnan(dqkv.float()).sum()))
        else:
            ms = timeit(lambda: L.call("passt_attn_bwd", L.ptr(qkv), L.ptr(out), L.ptr(dO), L.ptr(lse), L.ptr(dqkv),
                                       L.ptr(dbias), L.ptr(ws), B, N, H, scale, L.stream_ptr()))
            rec.update(ms=ms, tflops=10.0 * B * H * N * N * hd / ms / 1e9)
        log(**rec)


def g_gemm_2cta():
    """2-CTA kernel vs torch and vs the 1-CTA kernel (timing A/B)."""
    lib = L.load()
    torch.manual_seed(5)
    for (M, N, K) in [(256, 256, 64), (300, 512, 128), (1000, 768, 768), (30336, 2304, 768), (30336, 768, 3072),
                      (30336, 3072, 768), (30336, 768, 768)]:
        A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
        B = (torch.randn(N, K, device=dev) * 0.5).bfloat16()
        bias = torch.randn(N, device=dev)
        ref = A.float() @ B.float().t() + bias
        rec = dict(test="gemm2_tn", M=M, N=N, K=K)
        for two in (1, 0):
            lib.passt_gemm_set_2cta(two)
            C = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16)
            gemm(A, B, C, bias=bias, M=M, N=N, K=K, lda=K, ldb=K, ldc=N, mode=0)
            torch.cuda.synchronize()
            rec[f"relerr_{two}"] = relerr(C, ref)
            if M >= 30000:
                ms = timeit(lambda: gemm(A, B, C, bias=bias, M=M, N=N, K=K, lda=K, ldb=K, ldc=N, mode=0))
                rec[f"tflops_{two}"] = 2.0 * M * N * K / ms / 1e9
        log(**rec)
    for (Kt, M, N, splits) in [(64, 256, 256, 1), (1000, 768, 768, 4), (30336, 2304, 768, 16), (30336, 768, 3072, 8),
                               (30336, 3072, 768, 8), (30336, 768, 768, 24)]:
        A = (torch.randn(Kt, M, device=dev) * 0.5).bfloat16()
        B = (torch.randn(Kt, N, device=dev) * 0.5).bfloat16()
        ref = A.float().t() @ B.float()
        rec = dict(test="gemm2_wgrad", Kt=Kt, M=M, N=N, splits=splits)
        for two in (1, 0):
            lib.passt_gemm_set_2cta(two)
            C = torch.zeros(M, N, device=dev)
            gemm(A, B, C, M=M, N=N, K=Kt, lda=M, ldb=N, ldc=N, mode=4, splits=splits)
            torch.cuda.synchronize()
            rec[f"relerr_{two}"] = relerr(C, ref)
            if Kt >= 30000:
                ms = timeit(lambda: gemm(A, B, C, M=M, N=N, K=Kt, lda=M, ldb=N, ldc=N, mode=4, splits=splits))
                rec[f"tflops_{two}"] = 2.0 * M * N * Kt / ms / 1e9
        log(**rec)
    lib.passt_gemm_set_2cta(1)


def g_gemm_yardstick():
    """cuBLAS (torch.matmul) on the same shapes, as a yardstick for what the hardware sustains (not a product path)."""
    torch.manual_seed(5)
    lib = L.load()
    for (M, N, K) in [(30336, 2304, 768), (30336, 768, 3072), (30336, 3072, 768), (30336, 768, 768), (8192, 8192, 8192)]:
        A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
        B = (torch.randn(N, K, device=dev) * 0.5).bfloat16()
        C = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        rec = dict(test="yardstick", M=M, N=N, K=K)
        ms = timeit(lambda: torch.matmul(A, B.t(), out=C), iters=50)
        rec["cublas_tflops"] = 2.0 * M * N * K / ms / 1e9
        for two in (1, 0):
            lib.passt_gemm_set_2cta(two)
            ms = timeit(lambda: gemm(A, B, C, M=M, N=N, K=K, lda=K, ldb=K, ldc=N, mode=0), iters=50)
            rec[f"ours_tflops_{two}"] = 2.0 * M * N * K / ms / 1e9
        log(**rec)
    lib.passt_gemm_set_2cta(1)


def g_attn_timeline():
    """clock64 timeline of CTA 0 of the forward attention kernel (softmax thread / MMA lane / TMA producer)."""
    torch.manual_seed(4)
    B, N, H, hd = 64, 474, 12, 64
    C = H * hd
    qkv = torch.randn(B, N, 3 * C, device=dev).bfloat16()
    out = torch.empty(B, N, C, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, H, 512, device=dev)
    tl = torch.zeros(3 * 512, dtype=torch.int64, device=dev)
    for _ in range(3):
        L.call("passt_attn_fwd", L.ptr(qkv), L.ptr(out), L.ptr(lse), B, N, H, hd ** -0.5, L.stream_ptr())
    L.load().passt_attn_debug_timeline(L.ptr(tl))
    L.call("passt_attn_fwd", L.ptr(qkv), L.ptr(out), L.ptr(lse), B, N, H, hd ** -0.5, L.stream_ptr())
    torch.cuda.synchronize()
    L.load().passt_attn_debug_timeline(None)
    t = tl.cpu().tolist()
    sm, mm, pr = t[0:512], t[512:1024], t[1024:1536]
    t0 = min(x for x in sm + mm + pr if x > 0)
    rows = []
    for k in range(16):
        s = [x - t0 if x else None for x in sm[k * 5:k * 5 + 5]]
        m = [x - t0 if x else None for x in mm[k * 3:k * 3 + 3]]
        rows.append(dict(tile=k, softmax=s, mma=m, prod=(pr[k] - t0) if pr[k] else None))
    log(test="attn_timeline", legend="softmax: s_full|pass1|Odrain|pass2|arrive ; mma: p_full|PV issued|S next issued ; prod: kv slot free",
        rows=rows)


def g_attn_bwd_timeline():
    torch.manual_seed(4)
    B, N, H, hd = 64, 474, 12, 64
    C = H * hd
    qkv = torch.randn(B, N, 3 * C, device=dev).bfloat16()
    out = torch.empty(B, N, C, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, H, 512, device=dev)
    L.call("passt_attn_fwd", L.ptr(qkv), L.ptr(out), L.ptr(lse), B, N, H, hd ** -0.5, L.stream_ptr())
    dO = torch.randn(B, N, C, device=dev).bfloat16()
    dqkv = torch.empty(B, N, 3 * C, device=dev, dtype=torch.bfloat16)
    ws = torch.empty(L.load().passt_attn_bwd_workspace_bytes(B, N, H), dtype=torch.uint8, device=dev)
    tl = torch.zeros(2 * 512, dtype=torch.int64, device=dev)
    run = lambda: L.call("passt_attn_bwd", L.ptr(qkv), L.ptr(out), L.ptr(dO), L.ptr(lse), L.ptr(dqkv), None, L.ptr(ws), B,
                         N, H, hd ** -0.5, L.stream_ptr())
    for _ in range(3):
        run()
    L.load().passt_attn_bwd_debug_timeline(L.ptr(tl))
    run()
    torch.cuda.synchronize()
    L.load().passt_attn_bwd_debug_timeline(None)
    t = tl.cpu().tolist()
    cp, mm = t[0:512], t[512:1024]
    t0 = min(x for x in cp + mm if x > 0)
    rows = []
    for k in range(12):
        rows.append(dict(step=k, compute=[x - t0 if x else None for x in cp[k * 6:k * 6 + 6]],
                         mma=[x - t0 if x else None for x in mm[k * 4:k * 4 + 4]]))
    log(test="attn_bwd_timeline",
        legend="compute: step start|sdp_full|math done|pds arrived|dq_full|dq staged ; mma: qdo_full|S,dP issued|pds_full|dV,dK,dQ issued",
        rows=rows)


def g_gemm_bkn():
    """modes 0|16 and 3|16: B operand given as [K, N] row-major."""
    lib = L.load()
    torch.manual_seed(6)
    for (M, N, K) in [(300, 256, 64), (1000, 768, 3072), (30336, 768, 2304), (30336, 3072, 768)]:
        A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
        Bkn = (torch.randn(K, N, device=dev) * 0.5).bfloat16()
        aux = torch.randn(M, N, device=dev).bfloat16()
        ref = A.float() @ Bkn.float()
        rec = dict(test="gemm_bkn", M=M, N=N, K=K)
        for two in (1, 0):
            lib.passt_gemm_set_2cta(two)
            C = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16)
            gemm(A, Bkn, C, M=M, N=N, K=K, lda=K, ldb=N, ldc=N, mode=16)
            C3 = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16)
            cs = torch.zeros(N, device=dev)
            gemm(A, Bkn, C3, aux=aux, bias=cs, M=M, N=N, K=K, lda=K, ldb=N, ldc=N, mode=3 | 16, ld_aux=N)
            torch.cuda.synchronize()
            rec[f"m0_{two}"] = relerr(C, ref)
            rec[f"m3_{two}"] = relerr(C3, ref * aux.float())
            rec[f"cs_{two}"] = relerr(cs, (ref * aux.float()).sum(0))
            if M >= 30000:
                ms = timeit(lambda: gemm(A, Bkn, C, M=M, N=N, K=K, lda=K, ldb=N, ldc=N, mode=16))
                rec[f"tflops_{two}"] = 2.0 * M * N * K / ms / 1e9
        log(**rec)
    lib.passt_gemm_set_2cta(1)


GROUPS = {k[2:]: v for k, v in globals().items() if k.startswith("g_")}

if __name__ == "__main__":
    name = sys.argv[1]
    log(group=name, start=True, device=torch.cuda.get_device_name(0))
    try:
        GROUPS[name]()
        torch.cuda.synchronize()
        log(group=name, done=True)
    except Exception as e:  # noqa
        log(group=name, error=str(e), tb=traceback.format_exc())
        sys.exit(1)
