"""Timing of the HBM-bound kernels at cfg2 shapes with achieved GB/s (algorithmic bytes)."""
import sys; sys.path.insert(0, ".")
import torch
from nextgen_uia_b200 import ops
dev = torch.device("cuda:0")
B, N, D = 256, 197, 768
M = B * N
bf = torch.bfloat16
def tm(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it * 1e3
x = torch.randn(M, D, device=dev).to(bf); g = torch.randn(M, D, device=dev).to(bf); r = torch.randn(M, D, device=dev).to(bf)
w = torch.randn(D, device=dev); b = torch.randn(D, device=dev); ga = torch.randn(D, device=dev); gx = torch.randn(D, device=dev)
y, mean, rstd = ops.ln_fwd(x, w, b, 1e-6)
MB = M * D * 2 / 1e6
t = tm(lambda: ops.ln_fwd(x, w, b, 1e-6)); print(f"ln_fwd        {t:7.1f} us  {2*MB/t*1e-3*1e3:7.0f} GB/s")
t = tm(lambda: ops.ln_fwd(x, w, b, 1e-5, gamma=ga, gammax=gx)); print(f"ln_fwd(mix)   {t:7.1f} us  {2*MB/t*1e-3*1e3:7.0f} GB/s")
t = tm(lambda: ops.ln_bwd(g, x, mean, rstd, w, dres=r)); print(f"ln_bwd        {t:7.1f} us  {4*MB/t*1e-3*1e3:7.0f} GB/s")
z = lambda: torch.zeros(D, device=dev)
acc = [z() for _ in range(5)]
t = tm(lambda: ops.mona_pre_bwd(g, r, x, mean, rstd, w, b, ga, gx, *acc)); print(f"mona_pre_bwd  {t:7.1f} us  {4*MB/t*1e-3*1e3:7.0f} GB/s")
h = torch.randn(B, N, 64, device=dev).to(bf); dg = torch.randn(B, N, 64, device=dev).to(bf)
wts = [torch.randn(64, 1, 3, 3), torch.randn(64), torch.randn(64, 1, 5, 5), torch.randn(64), torch.randn(64, 1, 7, 7), torch.randn(64), torch.randn(64, 64, 1, 1) * 0.1, torch.randn(64)]
wts = [t_.to(dev) for t_ in wts]
grads = [torch.zeros_like(t_) for t_ in wts] + [torch.zeros(64, device=dev)]
hb = M * 64 * 2 / 1e6
t = tm(lambda: ops.mona_conv_fwd(h, wts, (14, 14), True, 0.1, 5)); print(f"mona_conv_fwd {t:7.1f} us  {2*hb/t*1e-3*1e3:7.0f} GB/s")
t = tm(lambda: ops.mona_conv_bwd(h, dg, wts, grads, (14, 14), True, 0.1, 5)); print(f"mona_conv_bwd {t:7.1f} us  {3*hb/t*1e-3*1e3:7.0f} GB/s")
img = torch.rand(B, 3, 224, 224, device=dev)
t = tm(lambda: ops.patchify(img, 16, bf)); print(f"patchify      {t:7.1f} us")
