"""InfoNCE core (fp32 CUDA-core logits, both LSEs, feature gradients) at the global batch sizes of 1..8 GPUs."""
import sys
sys.path.insert(0, ".")
import torch
from nextgen_uia_b200 import ops
dev = torch.device("cuda:0")
for Bg in (256, 512, 1024, 2048):
    i = torch.nn.functional.normalize(torch.randn(Bg, 512, device=dev), dim=1)
    t = torch.nn.functional.normalize(torch.randn(Bg, 512, device=dev), dim=1)
    f = lambda: ops.infonce_core(i, t, 0, 256, 0.07)
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): f()
    b.record(); torch.cuda.synchronize()
    print(f"Bg={Bg}: infonce_core {a.elapsed_time(b) / 10 * 1e3:.0f} us")
