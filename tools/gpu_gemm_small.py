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
x = torch.randn(M, 768, device=dev).to(bf); g = torch.randn(M, 64, device=dev).to(bf)
W1 = torch.randn(64, 768, device=dev).to(bf); W2 = torch.randn(768, 64, device=dev).to(bf)
b1 = torch.randn(64, device=dev); b2 = torch.randn(768, device=dev)
print("mona GEMM1  [M,768]x[64,768]^T + b      : %.1f us (ideal ~13 us: 84 MB)" % tm(lambda: ops.gemm(x, W1, bias=b1)))
print("mona GEMM2  [M,64]x[768,64]^T + b + res : %.1f us (ideal ~25 us: 161 MB)" % tm(lambda: ops.gemm(g, W2, bias=b2, aux=x, aux_mode=L.AUX_RESIDUAL)))
print("mona dg     [M,768]x[64,768]^T          : %.1f us" % tm(lambda: ops.gemm(x, W1)))
print("mona du     [M,64]x[768,64]^T           : %.1f us (ideal ~13 us: 84 MB)" % tm(lambda: ops.gemm(g, W2)))
for bn in (64, 128, 256):
    print(f"  GEMM2 block_n={bn}: %.1f us" % tm(lambda: ops.gemm(g, W2, bias=b2, aux=x, aux_mode=L.AUX_RESIDUAL, block_n=bn)))
xb = torch.randn(256 * 77, 768, device=dev).to(bf); Wq = torch.randn(2304, 768, device=dev).to(bf)
print("bert qkv [19712,768]x[2304,768]: %.1f us" % tm(lambda: ops.gemm(xb, Wq)))
