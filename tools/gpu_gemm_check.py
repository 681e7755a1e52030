"""Scratch GPU check for the tcgen05 GEMM (run under gpurun). Prints max errors and timings."""
import sys, time, ctypes, json
sys.path.insert(0, ".")
import torch
from nextgen_uia_b200 import _lib as L

lib = L.lib()
print("version", lib.ngu_version(), "selftest", lib.ngu_selftest_device(), flush=True)
dev = torch.device("cuda:0")
stream = torch.cuda.current_stream().cuda_stream


def gemm(A, B, bias=None, act=0, aux=None, aux_mode=0, save_pre=False, A2=None, B2=None, alpha=1.0, dtype=L.NGU_BF16, block_n=0):
    M, K = A.shape
    N = B.shape[0]
    C = torch.empty(M, N, device=dev, dtype=A.dtype)
    Pre = torch.empty(M, N, device=dev, dtype=A.dtype) if save_pre else None
    d = L.GemmDesc()
    d.A, d.lda = A.data_ptr(), A.stride(0)
    d.B, d.ldb = B.data_ptr(), B.stride(0)
    d.C, d.ldc = C.data_ptr(), C.stride(0)
    if A2 is not None:
        d.A2, d.lda2, d.B2, d.ldb2, d.K2 = A2.data_ptr(), A2.stride(0), B2.data_ptr(), B2.stride(0), A2.shape[1]
    d.bias = bias.data_ptr() if bias is not None else None
    if aux is not None:
        d.aux, d.ldaux = aux.data_ptr(), aux.stride(0)
    if save_pre:
        d.Pre, d.ldpre = Pre.data_ptr(), Pre.stride(0)
    d.M, d.N, d.K = M, N, K
    d.act, d.aux_mode, d.save_pre = act, aux_mode, int(save_pre)
    d.alpha = alpha
    d.dtype = dtype
    d.block_n = block_n
    L.check(lib.ngu_gemm(ctypes.byref(d), stream), "gemm")
    return C, Pre


def ref(A, B, bias=None, act=0, aux=None, aux_mode=0, A2=None, B2=None, alpha=1.0):
    y = A.float() @ B.float().t()
    if A2 is not None:
        y = y + A2.float() @ B2.float().t()
    y = y * alpha
    if bias is not None:
        y = y + bias
    pre = y
    if act:
        pp = y.detach().clone().requires_grad_(True)
        gg = torch.nn.functional.gelu(pp) if act == 1 else pp * torch.sigmoid(1.702 * pp)
        (pre,) = torch.autograd.grad(gg.sum(), pp)
    if aux_mode == 2:
        y = y * aux.float()
        return y, pre
    else:
        if act == 1:
            y = torch.nn.functional.gelu(y)
        elif act == 2:
            y = y * torch.sigmoid(1.702 * y)
        if aux_mode == 1:
            y = y + aux.float()
    return y, pre


def relerr(x, y):
    return ((x.float() - y.float()).abs().max() / y.float().abs().max().clamp_min(1e-6)).item()


torch.manual_seed(0)
results = []
cases = [
    # M, N, K, bn, kwargs
    (128, 256, 64, 0, {}),
    (128, 256, 768, 0, {}),
    (256, 512, 768, 0, {}),
    (1000, 768, 768, 0, dict(bias=True)),
    (1000, 2304, 768, 0, dict(bias=True)),
    (777, 3072, 768, 0, dict(bias=True, act=1, save_pre=True)),
    (777, 768, 3072, 0, dict(bias=True, aux_mode=1)),
    (777, 3072, 768, 0, dict(aux_mode=2)),
    (777, 3072, 768, 0, dict(bias=True, act=2)),
    (640, 64, 768, 0, dict(bias=True)),
    (640, 128, 768, 0, dict(bias=True)),
    (640, 512, 768, 128, dict()),
    (640, 2304, 768, 0, dict(bias=True, lora=16)),
    (640, 768, 768, 0, dict(bias=True, lora=8, aux_mode=1)),
    (300, 200, 264, 0, dict(bias=True)),
]
for (M, N, K, bn, kw) in cases:
    A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    B = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev) if kw.get("bias") else None
    aux = (torch.randn(M, N, device=dev)).bfloat16() if kw.get("aux_mode") else None
    A2 = B2 = None
    if kw.get("lora"):
        r = kw["lora"]
        A2 = (torch.randn(M, r, device=dev)).bfloat16()
        B2 = (torch.randn(N, r, device=dev) * 0.1).bfloat16()
    try:
        C, Pre = gemm(A, B, bias, kw.get("act", 0), aux, kw.get("aux_mode", 0), kw.get("save_pre", False), A2, B2, block_n=bn)
        torch.cuda.synchronize()
        R, Rpre = ref(A, B, bias, kw.get("act", 0), aux, kw.get("aux_mode", 0), A2, B2)
        e = relerr(C, R)
        ep = relerr(Pre, Rpre) if Pre is not None else None
        print(f"M={M} N={N} K={K} bn={bn} {kw}: relerr={e:.3e} pre={ep}", flush=True)
        results.append(dict(M=M, N=N, K=K, kw=str(kw), err=e, perr=ep))
    except Exception as ex:
        print(f"M={M} N={N} K={K} {kw}: EXC {ex}", flush=True)
        results.append(dict(M=M, N=N, K=K, kw=str(kw), exc=str(ex)))
        break

# SIMT fp32 check
A = torch.randn(300, 200, device=dev); B = torch.randn(150, 200, device=dev) * 0.1; bias = torch.randn(150, device=dev)
C, _ = gemm(A, B, bias, act=1, dtype=L.NGU_F32)
R, _ = ref(A, B, bias, act=1)
print("simt fp32 relerr", relerr(C, R), flush=True)

# perf at cfg2 shapes
def bench(M, N, K, iters=20, **kw):
    A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    B = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev)
    aux = torch.randn(M, N, device=dev).bfloat16() if kw.get("aux_mode") else None
    for _ in range(3):
        gemm(A, B, bias, kw.get("act", 0), aux, kw.get("aux_mode", 0), kw.get("save_pre", False))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        gemm(A, B, bias, kw.get("act", 0), aux, kw.get("aux_mode", 0), kw.get("save_pre", False))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    # cuBLAS comparator
    Bt = B.t().contiguous()
    for _ in range(3): torch.matmul(A, B.t())
    e0.record()
    for _ in range(iters): torch.matmul(A, B.t())
    e1.record(); torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / iters
    print(f"perf M={M} N={N} K={K} {kw}: {ms:.3f} ms {tf:.0f} TF/s | cublas {ms2:.3f} ms {2.0*M*N*K/ms2/1e9:.0f} TF/s", flush=True)
    results.append(dict(perf=(M, N, K), kw=str(kw), ms=ms, tflops=tf, cublas_ms=ms2))

M = 256 * 197
bench(M, 2304, 768)
bench(M, 768, 768, aux_mode=1)
bench(M, 3072, 768, act=1, save_pre=True)
bench(M, 768, 3072, aux_mode=1)
bench(M, 3072, 768, aux_mode=2)
bench(8192, 8192, 8192)
json.dump(results, open("gpurun_out/gemm_check.json", "w"), indent=1)
