"""Launch one tcgen05 GEMM shape a few times (ncu target)."""
import sys, ctypes
sys.path.insert(0, ".")
import torch
from nextgen_uia_b200 import _lib as L
lib = L.lib()
dev = torch.device("cuda:0")
M, N, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
act = int(sys.argv[4]) if len(sys.argv) > 4 else 0
A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
B = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
bias = torch.randn(N, device=dev)
C = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
d = L.GemmDesc()
d.A, d.lda, d.B, d.ldb, d.C, d.ldc = A.data_ptr(), K, B.data_ptr(), K, C.data_ptr(), N
d.bias = bias.data_ptr()
d.M, d.N, d.K = M, N, K
d.act = act
save = int(sys.argv[5]) if len(sys.argv) > 5 else 0
if save:
    Pre = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    d.Pre, d.ldpre, d.save_pre = Pre.data_ptr(), N, 1
d.alpha = 1.0
for _ in range(3):
    L.check(lib.ngu_gemm(ctypes.byref(d), torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
print("done")
