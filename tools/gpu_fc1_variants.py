import sys; sys.path.insert(0, ".")
import torch
from nextgen_uia_b200 import ops, _lib as L
dev = torch.device("cuda:0"); bf = torch.bfloat16
M = 256 * 197
def tm(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it * 1e3
x = (torch.randn(M, 768, device=dev) * 0.5).to(bf); W = (torch.randn(3072, 768, device=dev) * 0.05).to(bf); b = torch.randn(3072, device=dev)
aux = torch.randn(M, 3072, device=dev).to(bf)
print("plain            : %.0f us" % tm(lambda: ops.gemm(x, W, bias=b)))
print("gelu             : %.0f us" % tm(lambda: ops.gemm(x, W, bias=b, act=L.ACT_GELU)))
print("save_pre (no act): %.0f us" % tm(lambda: ops.gemm(x, W, bias=b, save_pre=True)))
print("gelu + save      : %.0f us" % tm(lambda: ops.gemm(x, W, bias=b, act=L.ACT_GELU, save_pre=True)))
print("dact (aux mul)   : %.0f us" % tm(lambda: ops.gemm(x, W, aux=aux, aux_mode=L.AUX_DACT)))
print("residual         : %.0f us" % tm(lambda: ops.gemm(x, W, bias=b, aux=aux, aux_mode=L.AUX_RESIDUAL)))
print("cublas           : %.0f us" % tm(lambda: torch.matmul(x, W.t())))
