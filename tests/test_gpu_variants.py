"""A/B switches of the kernel path: every variant must give the same results (through the C ABI).

  * attention forward: ping-pong kernel (attn_fwd2.cu, default) vs the two-CTAs-per-SM kernel (attn_fwd.cu) vs a torch
    fp32 softmax-attention reference, for token counts with even / odd / single query-tile counts;
  * programmatic dependent launch on / off;
  * residual adds fused into the proj / fc2 GEMM epilogues on / off;
  * attention-backward D = rowsum(dO o O) fused into the proj-dgrad GEMM epilogue on / off;
  * attention backward with eight vs sixteen compute warps.
"""
import pytest
import torch
import torch.nn.functional as F

from util import grad_metrics, quiet, relerr

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _attn_ref(qkv, H):
    B, N, C3 = qkv.shape
    C = C3 // 3
    q, k, v = qkv.float().reshape(B, N, 3, H, C // H).permute(2, 0, 3, 1, 4)
    att = ((q @ k.transpose(-2, -1)) * (C // H) ** -0.5).softmax(-1)
    return (att @ v).transpose(1, 2).reshape(B, N, C), torch.logsumexp((q @ k.transpose(-2, -1)) * (C // H) ** -0.5, -1)


@pytest.mark.parametrize("N,ramp", [(474, 0.0), (353, 0.0), (790, 0.0), (130, 0.0), (1190, 0.0), (64, 0.0), (474, 1.0), (790, 4.0)])
def test_attention_forward_variants_agree(N, ramp):
    """ramp > 0: key norms grow along the sequence, so the row maximum keeps rising far above the reference taken from the
    first keys -- exercises the lagged-maximum rescale (growth > 2^8) and the guarded redo path (> 2^64)."""
    from passt_b200 import _lib as L
    B, H = 3, 12
    C = H * 64
    torch.manual_seed(N)
    qkv = torch.randn(B, N, 3 * C, device=DEV) * 1.5
    if ramp:
        grow = 1.0 + ramp * torch.arange(N, device=DEV).float() / N * 6.0
        qkv[:, :, C:2 * C] *= grow.view(1, N, 1)          # keys
        qkv[:, :, :C] *= 3.0                              # queries
    qkv = qkv.bfloat16()
    npad = ((N + 127) // 128) * 128
    outs = {}
    for variant in (1, 2, 3, 4):
        L.load().passt_attn_fwd_set_variant(variant)
        o = torch.zeros(B, N, C, device=DEV, dtype=torch.bfloat16)
        lse = torch.zeros(B, H, npad, device=DEV)
        L.call("passt_attn_fwd", L.ptr(qkv), L.ptr(o), L.ptr(lse), B, N, H, 0.125, L.stream_ptr())
        torch.cuda.synchronize()
        outs[variant] = (o, lse)
    L.load().passt_attn_fwd_set_variant(2)
    ref_o, ref_lse = _attn_ref(qkv, H)
    for variant in (1, 2, 3, 4):
        o, lse = outs[variant]
        assert relerr(o, ref_o) < 1e-2, variant
        # log2-domain LSE of the scaled scores; pad rows are +inf
        got = lse[:, :, :N] * 0.6931471805599453
        assert (got - ref_lse).abs().max().item() < 2e-3 * max(1.0, ref_lse.abs().max().item() / 50), variant
        assert torch.isinf(lse[:, :, N:]).all()
    assert relerr(outs[2][0], outs[1][0]) < 4e-3          # same per-row arithmetic; bf16 output rounding at most
    assert torch.equal(outs[4][0], outs[2][0]) and torch.equal(outs[4][1], outs[2][1])   # same arithmetic, earlier MMA issue
    assert relerr(outs[3][0], outs[1][0]) < 8e-3          # a side-wide redo moves the reference of rows that did not need it: 1-2 bf16 ulps


@pytest.mark.parametrize("N", [474, 353, 130, 1190, 64])
def test_attention_backward_variants_agree(N):
    """attn_bwd_kernel (eight compute warps) vs attn_bwd2_kernel (sixteen, 32-query column quarters) vs torch autograd of
    the fp32 softmax attention on the same bf16 inputs."""
    from passt_b200 import _lib as L
    lib = L.load()
    B, H = 3, 12
    C = H * 64
    torch.manual_seed(N + 1)
    qkv = (torch.randn(B, N, 3 * C, device=DEV) * 1.2).bfloat16()
    dO = (torch.randn(B, N, C, device=DEV) * 0.5).bfloat16()
    npad = ((N + 127) // 128) * 128
    o = torch.zeros(B, N, C, device=DEV, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, npad, device=DEV)
    L.call("passt_attn_fwd", L.ptr(qkv), L.ptr(o), L.ptr(lse), B, N, H, 0.125, L.stream_ptr())
    ws = torch.empty(lib.passt_attn_bwd_workspace_bytes(B, N, H), dtype=torch.uint8, device=DEV)
    outs = {}
    try:
        for variant in (1, 2):
            lib.passt_attn_bwd_set_variant(variant)
            dqkv = torch.zeros_like(qkv)
            dbias = torch.zeros(3 * C, device=DEV)
            L.call("passt_attn_bwd", L.ptr(qkv), L.ptr(o), L.ptr(dO), L.ptr(lse), L.ptr(dqkv), L.ptr(dbias), L.ptr(ws), B, N,
                   H, 0.125, L.stream_ptr())
            torch.cuda.synchronize()
            outs[variant] = (dqkv, dbias)
    finally:
        lib.passt_attn_bwd_set_variant(1)
    x = qkv.float().requires_grad_(True)
    ref_o, _ = _attn_ref(x, H)
    ref_o.backward(dO.float())
    for variant in (1, 2):
        dqkv, dbias = outs[variant]
        for part, name in enumerate("qkv"):
            sl = slice(part * C, (part + 1) * C)
            assert relerr(dqkv[:, :, sl], x.grad[:, :, sl]) < 2e-2, (variant, name)
        assert relerr(dbias, x.grad.sum((0, 1))) < 2e-2, variant
    # same per-element arithmetic and the same MMA contraction order: identical up to the order of the fp32 dQ reduce-adds
    assert relerr(outs[2][0], outs[1][0]) < 4e-3
    assert relerr(outs[2][1], outs[1][1]) < 2e-3


def _small_train_net(seed=0):
    from passt_b200.passt import get_model, lighten_model
    torch.manual_seed(seed)
    with quiet():
        net = lighten_model(get_model(arch="passt_s_swa_p16_128_ap476", pretrained=False, s_patchout_t=40,
                                      s_patchout_f=4), cut_depth=9)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for p in net.parameters():
            p.add_(0.01 * torch.randn(p.shape, generator=g))
    return net.to(DEV).train()


def _run(net, x, y):
    net.zero_grad(set_to_none=True)
    torch.manual_seed(9)
    logits, _ = net(x)
    F.binary_cross_entropy_with_logits(logits, y).backward()
    torch.cuda.synchronize()
    return logits.detach().clone(), {k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None}


def test_pdl_on_off_same_results():
    from passt_b200 import _lib as L
    net = _small_train_net()
    torch.manual_seed(3)
    x = torch.randn(4, 1, 128, 1000, device=DEV)
    y = (torch.rand(4, 527, device=DEV) < 0.05).float()
    lib = L.load()
    was = lib.passt_get_pdl()
    try:
        lib.passt_set_pdl(1)
        la, ga = _run(net, x, y)
        la2, ga2 = _run(net, x, y)
        lib.passt_set_pdl(0)
        lb, gb = _run(net, x, y)
    finally:
        lib.passt_set_pdl(was)
    assert torch.equal(la, lb) and torch.equal(la, la2)
    for k in ga:
        # gradients use fp32 atomics / TMA reduce-adds whose order is not fixed: compare against the run-to-run spread
        assert relerr(gb[k], ga[k]) < 1e-4 + 2 * relerr(ga2[k], ga[k]), k


def test_fused_residual_and_dsum_switches():
    from passt_b200 import engine
    net = _small_train_net()
    torch.manual_seed(4)
    x = torch.randn(4, 1, 128, 1000, device=DEV)
    y = (torch.rand(4, 527, device=DEV) < 0.05).float()
    keep = (engine.FUSE_RESID, engine.FUSE_DSUM)
    try:
        engine.FUSE_RESID, engine.FUSE_DSUM = False, False
        l0, g0 = _run(net, x, y)
        engine.FUSE_RESID, engine.FUSE_DSUM = True, False
        l1, g1 = _run(net, x, y)
        engine.FUSE_RESID, engine.FUSE_DSUM = False, True
        l2, g2 = _run(net, x, y)
        engine.FUSE_RESID, engine.FUSE_DSUM = True, True
        l3, g3 = _run(net, x, y)
    finally:
        engine.FUSE_RESID, engine.FUSE_DSUM = keep
    # fused residual: the GEMM output joins the fp32 stream without a bf16 round trip -> tiny, bounded differences
    assert relerr(l1, l0) < 5e-3 and relerr(l3, l0) < 5e-3
    assert torch.equal(l2, l0)                                      # D-fusion only touches the backward
    for k in g0:
        m = grad_metrics(g2[k], g0[k])
        assert m["relmax"] < 2e-3 and m["cos"] > 0.99999, (k, m)     # same D up to fp32 summation order
        m = grad_metrics(g3[k], g0[k])
        assert m["relmax"] < 1e-2 and m["cos"] > 0.9995, (k, m)


@pytest.mark.parametrize("kw,T,mix", [(dict(s_patchout_t=40, s_patchout_f=4), 1000, False), (dict(u_patchout=400), 1000, True),
                                      (dict(), 1000, False), (dict(s_patchout_t=10, s_patchout_f=3), 500, True),
                                      (dict(s_patchout_t=40, s_patchout_f=4), 998, False)])
def test_single_kernel_patch_embed_matches_im2col_gemm(kw, T, mix):
    """passt_patch_embed (TMA gather + smem operand + tcgen05 GEMM + token table in one kernel) against the two-kernel
    path (passt_im2col + generic GEMM): same bf16 rounding of the patches, same fp32 accumulation -> logits agree to
    fp32-summation-order level, gradients (patch rows regathered in backward) too.  T = 998: TMA cannot describe the mel
    (rows not 16-byte aligned) and the engine must fall back silently to the two-kernel path."""
    from passt_b200 import engine
    from passt_b200.passt import get_model, lighten_model
    torch.manual_seed(0)
    with quiet():
        net = lighten_model(get_model(arch="passt_s_swa_p16_128_ap476", pretrained=False, **kw), cut_depth=10)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for p in net.parameters():
            p.add_(0.01 * torch.randn(p.shape, generator=g))
    net = net.to(DEV).train()
    B = 5
    torch.manual_seed(2)
    x = torch.randn(B, 1, 128, T, device=DEV)
    y = (torch.rand(B, 527, device=DEV) < 0.05).float()
    perm = torch.randperm(B).to(DEV)
    lam = (torch.rand(B) * 0.5 + 0.5).to(DEV)
    keep = engine.FUSE_PE
    res = {}
    try:
        for flag in (False, True):
            engine.FUSE_PE = flag
            if mix:
                net.fused_mixup(perm, lam)
            res[flag] = _run(net, x, y)
    finally:
        engine.FUSE_PE = keep
    (l0, g0), (l1, g1) = res[False], res[True]
    assert relerr(l1, l0) < 2e-4
    for k in g0:
        m = grad_metrics(g1[k], g0[k])
        assert m["relmax"] < 1e-3 and m["cos"] > 0.99999, (k, m)
