import sys; sys.path.insert(0, ".")
import torch
from nextgen_uia_b200 import ops
dev = torch.device("cuda:0"); bf = torch.bfloat16
B, N = 256, 197
h = torch.randn(B, N, 64, device=dev).to(bf); dg = torch.randn(B, N, 64, device=dev).to(bf)
wts = [torch.randn(64, 1, 3, 3), torch.randn(64), torch.randn(64, 1, 5, 5), torch.randn(64), torch.randn(64, 1, 7, 7), torch.randn(64), torch.randn(64, 64, 1, 1) * 0.1, torch.randn(64)]
wts = [t_.to(dev) for t_ in wts]
grads = [torch.zeros_like(t_) for t_ in wts] + [torch.zeros(64, device=dev)]
for _ in range(3):
    ops.mona_conv_fwd(h, wts, (14, 14), True, 0.1, 5)
    ops.mona_conv_bwd(h, dg, wts, grads, (14, 14), True, 0.1, 5)
torch.cuda.synchronize(); print("done")
