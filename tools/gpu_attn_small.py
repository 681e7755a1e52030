"""Forward attention at the BERT text-tower shape (B=256, N=77, H=12): per-tile kernel (impl 0) vs persistent (impl 2)."""
import sys
sys.path.insert(0, ".")
import torch
from nextgen_uia_b200 import ops
dev = torch.device("cuda:0")
for (B, N, H) in [(256, 77, 12), (256, 197, 12), (256, 128, 12), (256, 50, 12)]:
    D = H * 64
    qkv = torch.randn(B * N, 3 * D).to(dev, torch.bfloat16)
    for impl in (0, 2):
        f = lambda: ops.attn_fwd_packed(qkv, B, N, H, 64, impl=impl)
        for _ in range(3): f()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): f()
        b.record(); torch.cuda.synchronize()
        print(f"B={B} N={N} H={H} impl={impl}: {a.elapsed_time(b) / 20 * 1e3:.1f} us", flush=True)
