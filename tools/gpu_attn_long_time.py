"""Timing of the long-sequence attention kernels at the config-4 / config-5 shapes (CUDA events, warm)."""
import sys; sys.path.insert(0, ".")
import torch
from nextgen_uia_b200 import ops
dev = torch.device("cuda:0")
for (B, N, H) in ((64, 577, 16), (32, 485, 12)):
    D = H * 64
    qkv = torch.randn(B * N, 3 * D, device=dev).bfloat16()
    do = torch.randn(B * N, D, device=dev).bfloat16()
    for impl in (0, 1):
        o, lse = ops.attn_fwd_packed(qkv, B, N, H, 64, impl=impl)
        dq = ops.attn_bwd_packed(qkv, o, lse, do, B, N, H, 64, impl=impl)
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        for _ in range(3):
            o, lse = ops.attn_fwd_packed(qkv, B, N, H, 64, impl=impl)
        e[1].record()
        for _ in range(3):
            dq = ops.attn_bwd_packed(qkv, o, lse, do, B, N, H, 64, impl=impl)
        e[2].record()
        torch.cuda.synchronize()
        f, b = e[0].elapsed_time(e[1]) / 3, e[1].elapsed_time(e[2]) / 3
        fl = 4.0 * B * H * N * N * 64
        print(f"B={B} N={N} H={H} impl={'tcgen05' if impl == 0 else 'cuda-core'}: fwd {f*1e3:.0f} us ({fl/f/1e9:.0f} TFLOP/s)  bwd {b*1e3:.0f} us ({2.5*fl/b/1e9:.0f} TFLOP/s)")
