"""Per-phase cycle counts of the Mona conv-stage kernels (CTA 0), from a -DNGU_CONV_PROF build:
    make -C nextgen_uia_b200/csrc prof      # -> nextgen_uia_b200/libngu_b200_prof.so
    NGU_LIB=nextgen_uia_b200/libngu_b200_prof.so python tools/gpu_conv_phases.py
"""
import ctypes, sys
sys.path.insert(0, ".")
import torch
from nextgen_uia_b200 import ops, _lib as L
dev = torch.device("cuda:0"); bf = torch.bfloat16
B, N = 256, 197
h = torch.randn(B, N, 64, device=dev).to(bf); dg = torch.randn(B, N, 64, device=dev).to(bf)
wts = [torch.randn(64, 1, 3, 3), torch.randn(64), torch.randn(64, 1, 5, 5), torch.randn(64), torch.randn(64, 1, 7, 7), torch.randn(64), torch.randn(64, 64, 1, 1) * 0.1, torch.randn(64)]
wts = [t_.to(dev) for t_ in wts]
grads = [torch.zeros_like(t_) for t_ in wts] + [torch.zeros(64, device=dev)]
for p in (0.0, 0.1):
    for _ in range(3):
        ops.mona_conv_fwd(h, wts, (14, 14), True, p, 5)
        ops.mona_conv_bwd(h, dg, wts, grads, (14, 14), True, p, 5)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 32)()
    assert L.lib().ngu_debug_conv_prof(buf) == 0
    v = list(buf)
    names = ["load+weights", "cls+stencil z", "proj mma + gelu' -> da", "dP mma + dbp", "dz mma", "corr G", "param grads (atomics)", "dh stencil"]
    print(f"dropout p={p}: bwd total {v[8]-v[0]} cycles")
    for i, n in enumerate(names):
        print(f"  bwd {n:28s} {v[i+1]-v[i]:8d}")
    print(f"  fwd load+weights {v[17]-v[16]}, stencil {v[18]-v[17]}, proj+gelu+store {v[19]-v[18]}, total {v[19]-v[16]}")

# ---- fused Mona kernels (csrc/mona_fused.cu): CTA 0 timelines
from nextgen_uia_b200.adapters.mona import BaselineMona, BatchFirstMonaWrapper
m = BatchFirstMonaWrapper(BaselineMona(768, 64)).to(dev)
x = (torch.randn(B, 197, 768, device=dev) * 0.5).to(bf).requires_grad_(True)
gy = torch.randn(B, 197, 768, device=dev).to(bf)
for mode in ("eval", "train"):
    m.train(mode == "train")
    for _ in range(3):
        y = m(x, (14, 14)); y.backward(gy); x.grad = None
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 64)()
    assert L.lib().ngu_debug_fused_prof(buf) == 0
    v = list(buf)
    t0 = v[0]
    print(f"fused fwd stage ({mode}): kernel body {v[1]-v[0]} cycles (CTA 0, 2 images)")
    for it in range(2):
        b = 8 + it * 4
        print(f"  conv warps image {it}: start wait {v[b]-t0:7d}  got tile {v[b+1]-t0:7d}  stencil done {v[b+2]-t0:7d}  proj/gelu done {v[b+3]-t0:7d}")
        e = 24 + it * 3
        print(f"  stat warps image {it}: stats done {v[e]-t0:7d}  tmem+hs ready {v[e+1]-t0:7d}  epilogue done {v[e+2]-t0:7d}")
    names = ["load tile", "stencil z", "da (proj mma, gelu')", "dP mma", "dz mma", "corr G", "flush G", "dh stencil", "row phase", "final atomics"]
    print(f"fused bwd stage ({mode}): first image {v[41]-v[32]} cycles, kernel {v[42]-v[32]}")
    for i, n in enumerate(names):
        print(f"  {n:24s} {v[33+i]-v[32+i]:8d}")
