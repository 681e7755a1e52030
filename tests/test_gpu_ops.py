"""GPU parity of the individual kernels (through the C ABI) against the CPU oracle / golden vectors."""
import math

import pytest
import torch
import torch.nn.functional as F

from conftest import relerr, TOL, GTOL

pytestmark = pytest.mark.gpu
DTYPES = [torch.float32, torch.bfloat16]


def dev():
    return torch.device("cuda:0")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("M,N,K", [(300, 200, 264), (1000, 768, 768), (515, 3072, 768), (128, 64, 64), (77, 512, 640)])
def test_gemm_plain(dtype, M, N, K):
    from nextgen_uia_b200 import ops
    torch.manual_seed(0)
    A = (torch.randn(M, K) * 0.5).to(dev(), dtype)
    B = (torch.randn(N, K) * 0.05).to(dev(), dtype)
    bias = torch.randn(N, device=dev())
    C = ops.gemm(A, B, bias=bias)
    R = A.double().cpu() @ B.double().cpu().t() + bias.double().cpu()
    assert relerr(C, R) < TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
def test_gemm_epilogues(dtype):
    from nextgen_uia_b200 import ops, _lib as L
    torch.manual_seed(1)
    M, N, K, r = 640, 768, 256, 8
    A = (torch.randn(M, K) * 0.5).to(dev(), dtype)
    B = (torch.randn(N, K) * 0.1).to(dev(), dtype)
    bias = torch.randn(N, device=dev())
    aux = torch.randn(M, N).to(dev(), dtype)
    A2 = torch.randn(M, r).to(dev(), dtype)
    B2 = (torch.randn(N, r) * 0.1).to(dev(), dtype)
    base = (A.double() @ B.double().t() + bias.double()).cpu()
    # GELU + saved derivative gelu'(pre) (what the backward epilogue multiplies by)
    C, Der = ops.gemm(A, B, bias=bias, act=L.ACT_GELU, save_pre=True)
    bb = base.clone().requires_grad_(True)
    (dg_ref,) = torch.autograd.grad(F.gelu(bb).sum(), bb)
    def deq(D):   # bf16 path: the derivative is stored as one byte per element, q = round(d * 170 + 43)
        return (D.double().cpu() - 43.0) / 170.0 if D.dtype == torch.uint8 else D
    assert (Der.dtype == torch.uint8) == (dtype == torch.bfloat16)
    assert relerr(deq(Der), dg_ref) < TOL[dtype] and relerr(C, F.gelu(base)) < TOL[dtype]
    if dtype == torch.bfloat16:   # and the backward epilogue consumes that byte form: (acc) * act'(pre)
        Cb = ops.gemm(A, B, aux=Der, aux_mode=L.AUX_DACT)
        assert relerr(Cb, (base - bias.double().cpu()) * dg_ref) < TOL[dtype]
    # no activation: save_pre stores the pre-activation itself
    C0, P0 = ops.gemm(A, B, bias=bias, save_pre=True)
    assert relerr(P0, base) < TOL[dtype] and relerr(C0, base) < TOL[dtype]
    # QuickGELU (+ derivative)
    C, Der = ops.gemm(A, B, bias=bias, act=L.ACT_QUICKGELU, save_pre=True)
    bb = base.clone().requires_grad_(True)
    (dq_ref,) = torch.autograd.grad((bb * torch.sigmoid(1.702 * bb)).sum(), bb)
    assert relerr(C, base * torch.sigmoid(1.702 * base)) < TOL[dtype] and relerr(deq(Der), dq_ref) < TOL[dtype]
    C = ops.gemm(A, B, bias=bias, act=L.ACT_GELU)
    assert relerr(C, F.gelu(base)) < TOL[dtype]
    # residual
    C = ops.gemm(A, B, bias=bias, aux=aux, aux_mode=L.AUX_RESIDUAL)
    assert relerr(C, base + aux.double().cpu()) < TOL[dtype]
    # backward through the activation: (acc) * aux with aux = saved derivative
    C = ops.gemm(A, B, aux=aux, aux_mode=L.AUX_DACT)
    assert relerr(C, (base - bias.double().cpu()) * aux.double().cpu()) < TOL[dtype]
    # low-rank pair (LoRA) + alpha
    C = ops.gemm(A, B, bias=bias, A2=A2, B2=B2, alpha=0.5)
    ref = 0.5 * (A.double() @ B.double().t() + A2.double() @ B2.double().t()).cpu() + bias.double().cpu()
    assert relerr(C, ref) < TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("D", [768, 1024])
def test_layernorm(dtype, D):
    from nextgen_uia_b200 import ops
    torch.manual_seed(2)
    M = 333
    x = torch.randn(M, D) * 2 + 0.3
    w, b = 1 + 0.1 * torch.randn(D), 0.1 * torch.randn(D)
    g = torch.randn(M, D)
    dres = torch.randn(M, D)
    xd = x.to(dev(), dtype)
    y, mean, rstd = ops.ln_fwd(xd, w.to(dev()), b.to(dev()), 1e-6)
    xr = xd.double().cpu().requires_grad_(True)
    yr = F.layer_norm(xr, (D,), w.double(), b.double(), 1e-6)
    assert relerr(y, yr) < TOL[dtype]
    gd = g.to(dev(), dtype)
    (dxr,) = torch.autograd.grad((yr * gd.double().cpu()).sum(), xr)
    dx = ops.ln_bwd(gd, xd, mean, rstd, w.to(dev()), dres=dres.to(dev(), dtype))
    assert relerr(dx, dxr + dres.to(dtype).double()) < TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
def test_layernorm_strided_rows(dtype):
    """final norm on the CLS rows of [B,N,D] (rows = B, stride N*D)."""
    from nextgen_uia_b200 import ops
    torch.manual_seed(3)
    B, N, D = 5, 7, 768
    x = torch.randn(B, N, D).to(dev(), dtype)
    w, b = torch.randn(D).to(dev()), torch.randn(D).to(dev())
    y, mean, rstd = ops.ln_fwd(x, w, b, 1e-6, rows=B, ldx=N * D)
    ref = F.layer_norm(x[:, 0].double().cpu(), (D,), w.double().cpu(), b.double().cpu(), 1e-6)
    assert relerr(y, ref) < TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("B,N,H,causal", [(2, 197, 12, False), (3, 77, 12, False), (2, 50, 4, True), (3, 77, 8, True), (2, 200, 2, True)])
def test_attention(dtype, B, N, H, causal):
    from nextgen_uia_b200 import ops
    torch.manual_seed(4)
    dh = 64
    D = H * dh
    qkv = torch.randn(B * N, 3 * D).to(dev(), dtype)
    do = torch.randn(B * N, D).to(dev(), dtype)
    o, lse = ops.attn_fwd_packed(qkv, B, N, H, dh, causal=causal)
    dqkv = ops.attn_bwd_packed(qkv, o, lse, do, B, N, H, dh, causal=causal)
    t = qkv.double().cpu().requires_grad_(True)
    q, k, v = t.view(B, N, 3, H, dh).permute(2, 0, 3, 1, 4)
    ref = F.scaled_dot_product_attention(q, k, v, is_causal=causal).transpose(1, 2).reshape(B * N, D)
    (dref,) = torch.autograd.grad((ref * do.double().cpu()).sum(), t)
    assert relerr(o, ref) < TOL[dtype]
    assert relerr(dqkv, dref) < GTOL[dtype]


@pytest.mark.parametrize("B,N,H", [(30, 197, 12), (40, 77, 12), (70, 208, 5), (3, 193, 2), (2, 16, 1), (4, 65, 3), (2, 256, 3),
                                   (37, 128, 4)])
def test_attention_tc_persistent_paths(B, N, H):
    """tcgen05 attention against the CUDA-core kernels on shapes that exercise the persistent backward: several
    (batch, head) items per CTA (B*H > 148), one and two query tiles, trimmed second tiles, dead half-steps, both
    item parities of the alternating tile order.  impl = 0 is the pipelined persistent forward (attn_fwd_pipe_kernel), impl = 2 the
    one-CTA-per-query-tile forward."""
    from nextgen_uia_b200 import ops
    torch.manual_seed(11)
    dh = 64
    D = H * dh
    qkv = torch.randn(B * N, 3 * D).to(dev(), torch.bfloat16)
    do = torch.randn(B * N, D).to(dev(), torch.bfloat16)
    o, lse = ops.attn_fwd_packed(qkv, B, N, H, dh, impl=0)
    if N <= 640:                                     # the CUDA-core kernel's shared-memory score row ends there
        o1, lse1 = ops.attn_fwd_packed(qkv, B, N, H, dh, impl=1)
        assert relerr(o, o1) < 1e-2 and relerr(lse, lse1) < 1e-4
    o2, lse2 = ops.attn_fwd_packed(qkv, B, N, H, dh, impl=2)
    assert relerr(o2, o1) < 1e-2 and relerr(lse2, lse1) < 1e-4
    g = ops.attn_bwd_packed(qkv, o1, lse1, do, B, N, H, dh, impl=0)
    g1 = ops.attn_bwd_packed(qkv, o1, lse1, do, B, N, H, dh, impl=1)
    for sl in (slice(0, D), slice(D, 2 * D), slice(2 * D, 3 * D)):      # dq, dk, dv separately (different magnitudes)
        assert relerr(g[:, sl], g1[:, sl]) < 1e-2
    assert torch.isfinite(g.float()).all()


@pytest.mark.parametrize("dtype", DTYPES)
def test_wgrad_colsum(dtype):
    from nextgen_uia_b200 import ops
    torch.manual_seed(5)
    T, Mo, No = 3000, 768, 64
    X = torch.randn(T, Mo).to(dev(), dtype)
    Y = torch.randn(T, No).to(dev(), dtype)
    D = ops.wgrad(X, Y)
    ref = X.double().cpu().t() @ Y.double().cpu()
    assert relerr(D, ref) < TOL[dtype]
    D1 = ops.wgrad(X, Y, impl=1)                       # CUDA-core path
    assert relerr(D1, ref) < TOL[dtype]
    D2 = ops.wgrad(X, Y, out=D.clone())                # accumulates
    assert relerr(D2, 2 * ref) < TOL[dtype]
    if dtype == torch.bfloat16:                        # strided views (LoRA-padded factors), Mo = 2304, ragged T
        Xb = torch.randn(1111, 2304).to(dev(), dtype)
        Yb = torch.zeros(1111, 64, device=dev(), dtype=dtype); Yb[:, :8] = torch.randn(1111, 8).to(dev(), dtype)
        assert relerr(ops.wgrad(Xb, Yb), Xb.double().cpu().t() @ Yb.double().cpu()) < TOL[dtype]
    s = ops.colsum(X)
    assert relerr(s, X.double().cpu().sum(0)) < TOL[dtype]


@pytest.mark.parametrize("name", ["infonce_b8", "infonce_b37"])
@pytest.mark.parametrize("dtype", DTYPES)
def test_infonce_golden(golden, name, dtype):
    """loss and feature grads vs the reference InfoNCELoss outputs stored in tests/golden."""
    from src.losses import InfoNCELoss
    g = golden(name)
    I = g["I"].to(dev(), dtype).requires_grad_(True)
    T = g["T"].to(dev(), dtype).requires_grad_(True)
    loss = InfoNCELoss(temperature=g["temperature"])(I, T)
    loss.backward()
    # inputs were rounded to `dtype`; recompute the oracle on the rounded inputs for a like-for-like check
    from oracle import functional as OF
    Io, To = I.detach().double().cpu().requires_grad_(True), T.detach().double().cpu().requires_grad_(True)
    lo, _ = OF.info_nce(Io, To, g["temperature"])
    gI, gT = torch.autograd.grad(lo, [Io, To])
    # bf16 product path: the normalised features are rounded to bf16 for the tcgen05 logit GEMM (contract: 1e-2)
    assert abs(float(loss) - float(lo)) / abs(float(lo)) < (1e-4 if dtype == torch.float32 else 2e-3)
    assert relerr(I.grad, gI) < TOL[dtype] and relerr(T.grad, gT) < TOL[dtype]
    if dtype == torch.float32:
        assert abs(float(loss) - float(g["loss"])) / float(g["loss"]) < 1e-4
        assert relerr(I.grad, g["dI"]) < 1e-4 and relerr(T.grad, g["dT"]) < 1e-4


@pytest.mark.parametrize("Bg,Bl,r0", [(256, 256, 0), (1024, 256, 512), (2048, 256, 1792)])
def test_infonce_tensor_core_path_global_batch(Bg, Bl, r0):
    """bf16 product path of ngu_infonce_core at the global batch sizes of 1 / 4 / 8 GPUs: logits and both feature-gradient
    contractions on the tcgen05 GEMM; loss and the local rows' gradients vs the fp64 oracle on the same (gathered) features."""
    from nextgen_uia_b200 import ops
    from oracle import functional as OF
    torch.manual_seed(4)
    E = 512
    i = torch.nn.functional.normalize(torch.randn(Bg, E) + 0.5 * torch.randn(1, E), dim=1)
    t = torch.nn.functional.normalize(i + 0.7 * torch.randn(Bg, E), dim=1)       # correlated pairs: a non-trivial loss
    loss, di, dt = ops.infonce_core(i.to(dev()), t.to(dev()), r0, Bl, 0.07, tensor_cores=True)
    io, to = i.double().requires_grad_(True), t.double().requires_grad_(True)
    lo, _ = OF.info_nce(io, to, 0.07)         # inputs are unit rows: its normalisation is the identity up to 1e-7
    gi, gt = torch.autograd.grad(lo, [io, to])
    # d loss / d xhat of the oracle = gradient through its (identity) normalisation + the radial part it removes; compare the
    # tangential parts, which is what ngu_infonce_normalize_bwd keeps
    def tang(gr, x):
        return gr - x * (gr * x).sum(1, keepdim=True)
    assert abs(float(loss) - float(lo)) / abs(float(lo)) < 2e-3
    assert relerr(tang(di.double().cpu(), i.double()[r0:r0 + Bl]), gi[r0:r0 + Bl]) < 1e-2
    assert relerr(tang(dt.double().cpu(), t.double()[r0:r0 + Bl]), gt[r0:r0 + Bl]) < 1e-2


def test_fused_adamw_matches_torch():
    """clip_grad_norm_(1.0) + torch.optim.AdamW(betas=(0.9,0.95), wd=0.01) + CosineAnnealingLR, 5 updates."""
    from nextgen_uia_b200 import dp
    torch.manual_seed(0)
    shapes = [(64, 768), (768,), (64, 1, 7, 7), (3, 16, 1, 1)]
    ref = [torch.nn.Parameter(torch.randn(s)) for s in shapes]
    mine = [torch.nn.Parameter(p.detach().clone().to(dev())) for p in ref]
    opt = torch.optim.AdamW(ref, lr=1e-3, betas=(0.9, 0.95), weight_decay=0.01)
    sch = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=20, eta_min=1e-8)
    gb = dp.GradBuckets(mine, lambda p: 0, flatten_params=True)
    fo = dp.FusedAdamW(gb, 1e-3, (0.9, 0.95), 1e-8, 0.01, 1.0, 20, 1e-8)
    for it in range(5):
        gs = [torch.randn(s) * (3.0 if it % 2 == 0 else 0.01) for s in shapes]   # alternately clipped / not clipped
        for p, g_ in zip(ref, gs):
            p.grad = g_.clone()
        for p, g_ in zip(mine, gs):
            p.grad.copy_(g_.to(dev()))
        torch.nn.utils.clip_grad_norm_(ref, max_norm=1.0)
        opt.step(); sch.step()
        fo.step()
        for a, b in zip(mine, ref):
            assert relerr(a, b) < 1e-5
        assert float(gb.flat_grad.abs().sum()) == 0.0
    # non-finite loss: update skipped, grads still zeroed
    before = [p.detach().clone() for p in mine]
    for p in mine:
        p.grad.fill_(1.0)
    fo.step(loss=torch.full((1,), float("nan"), device=dev()))
    assert all(torch.equal(a, b) for a, b in zip(mine, before)) and float(gb.flat_grad.abs().sum()) == 0.0


def test_cast_batch_matches_single_casts():
    """ngu_cast_f32_batch (ops.CastPlan) == the per-tensor ngu_cast_f32 results, plain and transposed."""
    from nextgen_uia_b200 import ops
    torch.manual_seed(12)
    ps = [torch.nn.Parameter(torch.randn(r, c, device=dev())) for r, c in ((64, 768), (768, 64), (8, 768), (2304, 8))]
    plan = ops.CastPlan([(p, tr) for p in ps for tr in (False, True)], torch.bfloat16)
    plan.run()
    for p in ps:
        for tr in (False, True):
            assert torch.equal(plan.get(p, tr), ops.cast(p, torch.bfloat16, transpose=tr))
    with torch.no_grad():
        ps[0].mul_(2.0)
    plan.run()
    assert torch.equal(plan.get(ps[0], True), ops.cast(ps[0], torch.bfloat16, transpose=True))
    assert plan.valid() and plan.fresh
    with torch.no_grad():
        ps[1].add_(1.0)          # an in-place torch update makes the shadows stale until the next run()
    assert not plan.fresh
    plan.run()
    assert plan.fresh


@pytest.mark.parametrize("dtype", DTYPES)
def test_dropout_kernel_mask_statistics_and_regeneration(dtype):
    """ngu_dropout: keep rate ~ 1-p, kept values scaled by 1/(1-p), the same (seed, index) mask on every call (the
    backward regenerates it), a different mask for another seed, and `accumulate` adds into the output."""
    from nextgen_uia_b200 import ops
    torch.manual_seed(13)
    n, p = 1_000_003 - 3, 0.25            # multiple of 8, so every element goes through the vector path
    x = torch.ones(n, device=dev(), dtype=dtype)
    a = ops.dropout(x, p, 1234)
    b = ops.dropout(x, p, 1234)
    c = ops.dropout(x, p, 99)
    assert torch.equal(a, b) and not torch.equal(a, c)
    keep = (a != 0).float().mean().item()
    assert abs(keep - (1 - p)) < 3e-3
    kept = a[a != 0].float()
    assert torch.allclose(kept, torch.full_like(kept, 1 / (1 - p)), rtol=1e-2)
    # no structure across the four elements that share one hash
    m = (a != 0).float().view(-1, 4)
    assert (m.mean(0) - (1 - p)).abs().max().item() < 5e-3
    assert abs(torch.corrcoef(m.t())[0, 1].item()) < 1e-2
    out = torch.full_like(x, 2.0)
    ops.dropout(x, p, 1234, out=out, accumulate=True)
    assert torch.allclose(out.float(), a.float() + 2.0, rtol=1e-2)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("R,P", [(224, 16), (56, 14), (70, 14), (64, 8)])
def test_patchify_matches_unfold(dtype, R, P):
    """ngu_patchify == F.unfold(kernel = stride = P): the im2col of timm PatchEmbed / CLIP conv1, for the 16-pixel patches
    of ViT-B/16 (vector path) and the 14-pixel patches of ViT-L/14 (generic path, rows zero-padded to a multiple of 8)."""
    from nextgen_uia_b200 import ops
    torch.manual_seed(14)
    B = 3
    img = torch.rand(B, 3, R, R, device=dev())
    out = ops.patchify(img, P, dtype)
    G, K = R // P, 3 * P * P
    ref = F.unfold(img[:, :, :G * P, :G * P], kernel_size=P, stride=P).transpose(1, 2).reshape(B * G * G, K)
    assert out.shape == (B * G * G, (K + 7) // 8 * 8)
    assert relerr(out[:, :K], ref) < (1e-6 if dtype == torch.float32 else 4e-3)
    assert float(out[:, K:].float().abs().max()) == 0.0 if out.shape[1] > K else True


@pytest.mark.parametrize("N", [577, 485, 257, 640, 1024])
def test_attention_long_sequences_bf16(N):
    """Sequence lengths of configs 4 / 5 (ViT-L/14@336: 577 tokens, ViT-B/16@352: 485) and the edges of the range: the
    key-tiled tcgen05 kernels of attention_long.cu (online softmax forward, FlashAttention-2-style dQ and dK/dV backward).
    Output, LSE and all three gradients vs SDPA in fp64; the CUDA-core kernels (impl = 1) must agree as well, which also
    proves the default dispatch took a different (the tensor-core) path."""
    from nextgen_uia_b200 import ops, _lib as L
    torch.manual_seed(15)
    B, H, dh = 2, 3, 64
    D = H * dh
    qkv = torch.randn(B * N, 3 * D).to(dev(), torch.bfloat16)
    do = torch.randn(B * N, D).to(dev(), torch.bfloat16)
    n0 = L.launch_count()
    o, lse = ops.attn_fwd_packed(qkv, B, N, H, dh)
    dqkv = ops.attn_bwd_packed(qkv, o, lse, do, B, N, H, dh)
    assert L.launch_count() - n0 == 4          # fwd + (delta, dQ, dK/dV): the tiled tcgen05 path, not the 2-launch CUDA-core one
    t = qkv.double().cpu().requires_grad_(True)
    q, k, v = t.view(B, N, 3, H, dh).permute(2, 0, 3, 1, 4)
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * N, D)
    (dref,) = torch.autograd.grad((ref * do.double().cpu()).sum(), t)
    lse_ref = torch.logsumexp((q @ k.transpose(-1, -2)) * dh ** -0.5, -1)          # [B, H, N]
    assert relerr(o, ref) < 1e-2
    assert float((lse.double().cpu() - lse_ref.detach()).abs().max()) < 2e-2
    for sl in (slice(0, D), slice(D, 2 * D), slice(2 * D, 3 * D)):
        assert relerr(dqkv[:, sl], dref[:, sl]) < 1.5e-2
    if N <= 640:                                     # the CUDA-core kernel's shared-memory score row ends there
        o1, lse1 = ops.attn_fwd_packed(qkv, B, N, H, dh, impl=1)
        assert relerr(o, o1) < 1e-2


@pytest.mark.parametrize("B,N,H,causal", [(40, 197, 12, False), (13, 256, 12, True), (50, 77, 8, True), (21, 128, 8, False), (16, 224, 10, False)])
def test_attention_pipe_forward_masks_many_items(B, N, H, causal):
    """The pipelined forward with several (batch, head) items per CTA (B*H > 148) and a mask: per-batch key-padding lengths
    (non-causal cases) or the causal mask, against the CUDA-core kernel.  Covers both run lengths (N <= 224 and N = 256), one and
    two query tiles, the four K/V buffers of the single-tile case and dead row quarters of a short second tile."""
    from nextgen_uia_b200 import ops, _lib as L
    torch.manual_seed(23)
    dh = 64
    D = H * dh
    qkv = torch.randn(B * N, 3 * D).to(dev(), torch.bfloat16)
    lens = None
    if not causal:
        lens = torch.randint(1, N + 1, (B,))
        lens[0], lens[-1] = N, 1
        lens = lens.to(dev(), torch.int32)
    n0 = L.launch_count()
    o, lse = ops.attn_fwd_packed(qkv, B, N, H, dh, causal=causal, kv_len=lens, impl=0)
    assert L.launch_count() - n0 == 1
    o1, lse1 = ops.attn_fwd_packed(qkv, B, N, H, dh, causal=causal, kv_len=lens, impl=1)
    assert torch.isfinite(o.float()).all()
    assert relerr(o, o1) < 1e-2 and relerr(lse, lse1) < 1e-4


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("B,N,H", [(5, 77, 12), (3, 197, 2), (4, 256, 1)])
def test_attention_forward_key_padding_lengths(dtype, B, N, H):
    """ngu_attn_fwd with kv_len: keys at positions >= kv_len[b] are masked (right-padded batches).  Valid query rows
    match SDPA with the boolean key mask; the backward entry point refuses kv_len."""
    from nextgen_uia_b200 import ops, _lib as L
    torch.manual_seed(16)
    dh = 64
    D = H * dh
    qkv = torch.randn(B * N, 3 * D).to(dev(), dtype)
    lens = torch.randint(1, N + 1, (B,))
    lens[0] = N
    lens[-1] = 1
    o, lse = ops.attn_fwd_packed(qkv, B, N, H, dh, kv_len=lens.to(dev(), torch.int32))
    t = qkv.double().cpu()
    q, k, v = t.view(B, N, 3, H, dh).permute(2, 0, 3, 1, 4)
    mask = (torch.arange(N)[None, :] < lens[:, None])[:, None, None, :]
    ref = F.scaled_dot_product_attention(q, k, v, attn_mask=mask).transpose(1, 2).reshape(B, N, D)
    got = o.view(B, N, D)
    for b in range(B):
        assert relerr(got[b, :lens[b]], ref[b, :lens[b]]) < TOL[dtype]
    # backward with the same mask: gradients of valid positions vs autograd through masked SDPA, zero for masked keys.
    # The cotangent of padded query rows is zero (nothing downstream reads them).
    do = torch.randn(B, N, D)
    for b in range(B):
        do[b, lens[b]:] = 0
    dqkv = ops.attn_bwd_packed(qkv, o, lse, do.view(B * N, D).to(dev(), dtype), B, N, H, dh, kv_len=lens.to(dev(), torch.int32))
    tt = qkv.double().cpu().requires_grad_(True)
    q2, k2, v2 = tt.view(B, N, 3, H, dh).permute(2, 0, 3, 1, 4)
    ref2 = F.scaled_dot_product_attention(q2, k2, v2, attn_mask=mask).transpose(1, 2).reshape(B, N, D)
    (dref,) = torch.autograd.grad((ref2 * do.double()).sum(), tt)
    got_g, ref_g = dqkv.view(B, N, 3 * D), dref.view(B, N, 3 * D)
    for b in range(B):
        for sl in (slice(0, D), slice(D, 2 * D), slice(2 * D, 3 * D)):
            a_, r_ = got_g[b, :lens[b], sl].double().cpu(), ref_g[b, :lens[b], sl]
            # (a single valid key makes dq and dk vanish identically: scale by at least the cotangent's size)
            assert float((a_ - r_).abs().max()) / max(float(r_.abs().max()), 1.0) < GTOL[dtype]
        if lens[b] < N:
            assert float(got_g[b, lens[b]:, D:].float().abs().max()) == 0.0      # dk, dv of masked keys


def test_entry_points_are_cuda_graph_capturable():
    """The C-ABI promises stream-ordered, sync-free, allocation-free launches: a GEMM + LayerNorm + attention (fwd, bwd) +
    Mona conv sequence is captured into a CUDA graph (after one eager warm-up that sets the kernels' attributes) and
    replayed on fresh inputs; results equal the eager ones bit for bit."""
    from nextgen_uia_b200 import ops
    torch.manual_seed(17)
    B, N, H, dh = 4, 197, 12, 64
    D = H * dh
    bf = torch.bfloat16
    x = torch.randn(B * N, D, device=dev()).to(bf)
    Wq = (torch.randn(3 * D, D, device=dev()) * 0.03).to(bf)
    bq = torch.randn(3 * D, device=dev())
    w, b = torch.randn(D, device=dev()), torch.randn(D, device=dev())
    do = torch.randn(B * N, D, device=dev()).to(bf)

    def run():
        xn, _, _ = ops.ln_fwd(x, w, b, 1e-6)
        qkv = ops.gemm(xn, Wq, bias=bq)
        o, lse = ops.attn_fwd_packed(qkv, B, N, H, dh)
        g = ops.attn_bwd_packed(qkv, o, lse, do, B, N, H, dh)
        return o, g

    o0, g0 = run()                      # warm-up: function attributes, tensor-map encoder lookup
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        run()
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        o1, g1 = run()
    x.copy_(torch.randn(B * N, D, device=dev()).to(bf))          # new input in the captured buffer
    graph.replay()
    torch.cuda.synchronize()
    o2, g2 = run()
    assert torch.equal(o1, o2) and torch.equal(g1, g2)
    assert not torch.equal(o1, o0)


@pytest.mark.parametrize("M,N,K", [(50432, 768, 64), (1000, 768, 64), (300, 800, 128), (4096, 1024, 64)])
def test_gemm_residual_tma_aux_ring(M, N, K, monkeypatch):
    """Short-K GEMM + residual (Mona project2): the elementwise operand reaches the epilogue through per-warp TMA rings;
    ragged M / N (zero-filled slabs, clipped stores) and agreement with the per-lane-load epilogue (NGU_GEMM_AUXTMA is read
    once per process, so the A/B here is against the fp64 reference)."""
    from nextgen_uia_b200 import ops, _lib as L
    torch.manual_seed(7)
    A = torch.randn(M, K).to(dev(), torch.bfloat16)
    W = (torch.randn(N, K) * 0.1).to(dev(), torch.bfloat16)
    R = torch.randn(M, N).to(dev(), torch.bfloat16)
    b = torch.randn(N, device=dev())
    y = ops.gemm(A, W, bias=b, aux=R, aux_mode=L.AUX_RESIDUAL)
    ref = A.double().cpu() @ W.double().cpu().t() + b.double().cpu() + R.double().cpu()
    assert relerr(y, ref) < 1e-2


@pytest.mark.parametrize("M", [50432, 777])
def test_gemm_mona_dx_epilogue(M):
    """NGU_AUX_MONA_DX: C = A B^T + aux + beta_r * aux2 + alpha_r (two operands through the TMA rings, per-row scalars)."""
    from nextgen_uia_b200 import ops, _lib as L
    torch.manual_seed(8)
    N, K = 768, 128
    A = torch.randn(M, K).to(dev(), torch.bfloat16)
    W = (torch.randn(N, K) * 0.1).to(dev(), torch.bfloat16)
    dy = torch.randn(M, N).to(dev(), torch.bfloat16)
    x = torch.randn(M, N).to(dev(), torch.bfloat16)
    rowab = torch.randn(M, 2, device=dev())
    y = ops.gemm(A, W, aux=dy, aux_mode=L.AUX_MONA_DX, aux2=x, rowab=rowab)
    ab = rowab.double().cpu()
    ref = A.double().cpu() @ W.double().cpu().t() + dy.double().cpu() + ab[:, 1:2] * x.double().cpu() + ab[:, 0:1]
    assert relerr(y, ref) < 1e-2
