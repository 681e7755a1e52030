"""InfoNCE core (logits, both LSEs, feature gradients; CUDA-core fp32 check mode vs the tcgen05 bf16 product path) at the global batch sizes of 1..8 GPUs."""
import sys
sys.path.insert(0, ".")
import torch
from nextgen_uia_b200 import ops
dev = torch.device("cuda:0")
for Bg in (256, 512, 1024, 2048):
    i = torch.nn.functional.normalize(torch.randn(Bg, 512, device=dev), dim=1)
    t = torch.nn.functional.normalize(torch.randn(Bg, 512, device=dev), dim=1)
    for tcs in (False, True):
        f = lambda: ops.infonce_core(i, t, 0, 256, 0.07, tensor_cores=tcs)
        for _ in range(3): f()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            f()
        g.replay(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): g.replay()
        b.record(); torch.cuda.synchronize()
        print(f"Bg={Bg}: infonce_core ({'tcgen05' if tcs else 'cuda cores'}, graph replay) {a.elapsed_time(b) / 10 * 1e3:.0f} us")
