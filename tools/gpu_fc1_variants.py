"""fc1-shaped GEMM (M = 50432, N = 3072, K = 768) epilogue variants x tile variants (auto / CTA pair / multicast cluster / single CTA)."""
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
_, der = ops.gemm(x, W, bias=b, act=L.ACT_GELU, save_pre=True)
print("derivative dtype", der.dtype)
rows = (("auto", 0),) if len(sys.argv) > 1 else (("auto", 0), ("pair", 2256), ("cluster", 256), ("single", 1256))
for name, bn in rows:
    r = []
    r.append(("plain", tm(lambda: ops.gemm(x, W, bias=b, block_n=bn))))
    r.append(("gelu", tm(lambda: ops.gemm(x, W, bias=b, act=L.ACT_GELU, block_n=bn))))
    r.append(("gelu+save(u8)", tm(lambda: ops.gemm(x, W, bias=b, act=L.ACT_GELU, save_pre=True, block_n=bn))))
    r.append(("qgelu+save(u8)", tm(lambda: ops.gemm(x, W, bias=b, act=L.ACT_QUICKGELU, save_pre=True, block_n=bn))))
    r.append(("dact bf16", tm(lambda: ops.gemm(x, W, aux=aux, aux_mode=L.AUX_DACT, block_n=bn))))
    r.append(("dact u8", tm(lambda: ops.gemm(x, W, aux=der, aux_mode=L.AUX_DACT, block_n=bn))))
    r.append(("residual", tm(lambda: ops.gemm(x, W, bias=b, aux=aux, aux_mode=L.AUX_RESIDUAL, block_n=bn))))
    print(f"{name:8s}: " + "  ".join(f"{k} {v:.0f}" for k, v in r))
print("cublas           : %.0f us" % tm(lambda: torch.matmul(x, W.t())))
