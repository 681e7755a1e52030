"""CTA-pair (cta_group::2) GEMM variant: correctness against torch and timing against the default variant.
Usage: python tools/gpu_gemm_pair.py   (run under a short `timeout`)"""
import sys, ctypes
sys.path.insert(0, ".")
import torch
from nextgen_uia_b200 import _lib as L
lib = L.lib()
dev = torch.device("cuda:0")


def gemm(A, B, bias=None, act=0, aux=None, aux_mode=0, save_pre=False, A2=None, B2=None, block_n=0, C=None, Pre=None):
    M, K = A.shape
    N = B.shape[0]
    C = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if C is None else C
    d = L.GemmDesc()
    d.A, d.lda, d.B, d.ldb, d.C, d.ldc = A.data_ptr(), K, B.data_ptr(), K, C.data_ptr(), N
    if bias is not None: d.bias = bias.data_ptr()
    if aux is not None: d.aux, d.ldaux, d.aux_mode = aux.data_ptr(), N, aux_mode
    if A2 is not None:
        d.A2, d.lda2, d.B2, d.ldb2, d.K2 = A2.data_ptr(), A2.shape[1], B2.data_ptr(), B2.shape[1], A2.shape[1]
    d.M, d.N, d.K = M, N, K
    d.act = act
    if save_pre:
        Pre = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if Pre is None else Pre
        d.Pre, d.ldpre, d.save_pre = Pre.data_ptr(), N, 1
    d.alpha = 1.0
    d.block_n = block_n
    L.check(lib.ngu_gemm(ctypes.byref(d), torch.cuda.current_stream().cuda_stream))
    return C, Pre


def relerr(x, y):
    return ((x.float() - y.float()).abs().max() / y.float().abs().max().clamp_min(1e-6)).item()


torch.manual_seed(0)
ok = True
for (M, N, K, kw) in [(256, 256, 64, {}), (256, 256, 768, {}), (1000, 768, 768, dict(bias=1)), (777, 2304, 768, dict(bias=1, lora=64)),
                      (777, 3072, 768, dict(bias=1, act=1, save_pre=1)), (1300, 768, 3072, dict(bias=1, aux_mode=1)),
                      (640, 512, 264, dict(bias=1)), (50432, 768, 768, dict(bias=1, aux_mode=1))]:
    A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    B = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev) if kw.get("bias") else None
    aux = torch.randn(M, N, device=dev).bfloat16() if kw.get("aux_mode") else None
    A2 = B2 = None
    if kw.get("lora"):
        A2 = torch.randn(M, kw["lora"], device=dev).bfloat16(); B2 = (torch.randn(N, kw["lora"], device=dev) * 0.1).bfloat16()
    for bn in (2256, 2128):
        C1, P1 = gemm(A, B, bias, kw.get("act", 0), aux, kw.get("aux_mode", 0), bool(kw.get("save_pre")), A2, B2, block_n=bn)
        C0, P0 = gemm(A, B, bias, kw.get("act", 0), aux, kw.get("aux_mode", 0), bool(kw.get("save_pre")), A2, B2, block_n=0)
        torch.cuda.synchronize()
        R = A.float() @ B.float().t()
        if A2 is not None: R += A2.float() @ B2.float().t()
        if bias is not None: R += bias
        same = torch.equal(C1, C0) and (P1 is None or torch.equal(P1, P0))
        e = relerr(C0 if kw.get("act") or kw.get("aux_mode") else C1, C0) if kw.get("act") or kw.get("aux_mode") else relerr(C1, R)
        print(f"M={M} N={N} K={K} bn={bn} {kw}: bit-identical to default={same} relerr={e:.2e}", flush=True)
        ok &= same
print("PAIR_OK" if ok else "PAIR_MISMATCH", flush=True)


def timeit(f, iters=30):
    for _ in range(5): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


M = 50432
for (N, K, kw) in [(2304, 768, dict(bias=1)), (768, 768, dict(bias=1, aux_mode=1)), (3072, 768, dict(bias=1, act=1, save_pre=1)),
                   (768, 3072, dict(bias=1, aux_mode=1)), (3072, 768, dict(aux_mode=2)), (768, 832, dict(bias=1, aux_mode=1)), (2304, 832, dict(bias=1))]:
    A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    B = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev) if kw.get("bias") else None
    aux = torch.randn(M, N, device=dev).bfloat16() if kw.get("aux_mode") else None
    C = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    P = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    line = f"perf N={N} K={K} {kw}:"
    for bn in (0, 1000, 2256):
        us = timeit(lambda: gemm(A, B, bias, kw.get("act", 0), aux, kw.get("aux_mode", 0), bool(kw.get("save_pre")), block_n=bn, C=C, Pre=P))
        line += f"  bn={bn}: {us:.1f} us {2.0*M*N*K/us/1e6:.0f} TF/s"
    print(line, flush=True)
