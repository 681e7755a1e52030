"""A few fused Mona forward + backward passes at the benchmark shape (for ncu -k captures)."""
import sys; sys.path.insert(0, ".")
import torch
from nextgen_uia_b200.adapters.mona import BaselineMona, BatchFirstMonaWrapper
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = BatchFirstMonaWrapper(BaselineMona(768, 64)).to(dev).train()
x = (torch.randn(256, 197, 768, device=dev) * 0.5).bfloat16().requires_grad_(True)
g = torch.randn(256, 197, 768, device=dev).bfloat16()
for _ in range(4):
    y = m(x, (14, 14)); y.backward(g); x.grad = None
torch.cuda.synchronize(); print("done")
