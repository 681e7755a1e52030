"""A/B timing of the GEMM step shapes: libngu_b200.so (new) against libngu_b200_old.so (previous build), interleaved."""
import sys, ctypes
sys.path.insert(0, ".")
import torch
from nextgen_uia_b200 import _lib as L
new = L.lib()
old = ctypes.CDLL("nextgen_uia_b200/libngu_b200_old.so")
old.ngu_gemm.argtypes = new.ngu_gemm.argtypes
old.ngu_gemm.restype = ctypes.c_int
dev = torch.device("cuda:0")


def desc(A, B, C, bias=None, act=0, aux=None, aux_mode=0, Pre=None):
    M, K = A.shape
    N = B.shape[0]
    d = L.GemmDesc()
    d.A, d.lda, d.B, d.ldb, d.C, d.ldc = A.data_ptr(), K, B.data_ptr(), K, C.data_ptr(), N
    if bias is not None: d.bias = bias.data_ptr()
    if aux is not None: d.aux, d.ldaux, d.aux_mode = aux.data_ptr(), N, aux_mode
    d.M, d.N, d.K = M, N, K
    d.act = act
    if Pre is not None: d.Pre, d.ldpre, d.save_pre = Pre.data_ptr(), N, 1
    d.alpha = 1.0
    return d


def timeit(lib, d, iters=20):
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3): lib.ngu_gemm(ctypes.byref(d), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): lib.ngu_gemm(ctypes.byref(d), st)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


M = int(sys.argv[1]) if len(sys.argv) > 1 else 50432
for (N, K, kw) in [(2304, 768, dict(bias=1)), (768, 768, dict(bias=1, aux_mode=1)), (3072, 768, dict(bias=1, act=1, save_pre=1)),
                   (3072, 768, dict(bias=1, act=1)), (768, 3072, dict(bias=1, aux_mode=1)), (3072, 768, dict(aux_mode=2)), (768, 2304, dict()),
                   (768, 768, dict()), (768, 3072, dict())]:
    A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    B = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev) if kw.get("bias") else None
    aux = torch.randn(M, N, device=dev).bfloat16() if kw.get("aux_mode") else None
    C = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    P = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if kw.get("save_pre") else None
    d = desc(A, B, C, bias, kw.get("act", 0), aux, kw.get("aux_mode", 0), P)
    t = [[], []]
    for rep in range(3):
        t[0].append(timeit(old, d)); t[1].append(timeit(new, d))
    print(f"M={M} N={N} K={K} {kw}: old {min(t[0]):.1f} us  new {min(t[1]):.1f} us  ({min(t[0]) / min(t[1]):.3f}x)", flush=True)
