"""One DACT-mode GEMM launch at the fc2-backward shape (ncu target)."""
import sys, ctypes
sys.path.insert(0, ".")
import torch
from nextgen_uia_b200 import _lib as L
lib = L.lib()
dev = torch.device("cuda:0")
M, N, K = 50432, 3072, 768
A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
B = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
aux = torch.randn(M, N, device=dev).bfloat16()
C = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
d = L.GemmDesc()
d.A, d.lda, d.B, d.ldb, d.C, d.ldc = A.data_ptr(), K, B.data_ptr(), K, C.data_ptr(), N
d.aux, d.ldaux, d.aux_mode = aux.data_ptr(), N, 2
d.M, d.N, d.K = M, N, K
d.alpha = 1.0
for _ in range(3):
    L.check(lib.ngu_gemm(ctypes.byref(d), torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
print("done")
