"""ncu target: one launch each of the block's epilogue-heavy GEMMs (fc1 plain / GELU / GELU + derivative byte, fc2 dgrad * derivative byte,
proj + residual) after a warm-up launch of each."""
import sys; sys.path.insert(0, ".")
import torch
from nextgen_uia_b200 import ops, _lib as L
dev = torch.device("cuda:0"); bf = torch.bfloat16
M = 256 * 197
x = (torch.randn(M, 768, device=dev) * 0.5).to(bf); W = (torch.randn(3072, 768, device=dev) * 0.05).to(bf); b = torch.randn(3072, device=dev)
Wp = (torch.randn(768, 768, device=dev) * 0.05).to(bf); bp = torch.randn(768, device=dev)
_, der = ops.gemm(x, W, bias=b, act=L.ACT_GELU, save_pre=True)
for rep in range(2):
    if rep == 1:
        torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStart()
    ops.gemm(x, W, bias=b)
    ops.gemm(x, W, bias=b, act=L.ACT_GELU)
    ops.gemm(x, W, bias=b, act=L.ACT_GELU, save_pre=True)
    ops.gemm(x, W, aux=der, aux_mode=L.AUX_DACT)
    ops.gemm(x, Wp, bias=bp, aux=x, aux_mode=L.AUX_RESIDUAL)
    ops.gemm(x, Wp, bias=bp)
torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStop()
print("done")
