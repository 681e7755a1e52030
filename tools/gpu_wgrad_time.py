import sys; sys.path.insert(0, ".")
import torch
from nextgen_uia_b200 import ops
dev = torch.device("cuda:0")
T = 256 * 197
X = torch.randn(T, 768, device=dev).bfloat16(); Y = torch.randn(T, 64, device=dev).bfloat16()
D = torch.zeros(768, 64, device=dev)
def tm(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it * 1e3
print("wgrad tc 768x64: %.1f us  simt: %.1f us" % (tm(lambda: ops.wgrad(X, Y, out=D)), tm(lambda: ops.wgrad(X, Y, out=D, impl=1))))
X2 = torch.randn(T, 2304, device=dev).bfloat16(); D2 = torch.zeros(2304, 64, device=dev)
print("wgrad tc 2304x64: %.1f us" % tm(lambda: ops.wgrad(X2, Y, out=D2)))
