"""Warm CUDA-event timing of the tcgen05 attention forward / backward at the ViT-B/16 (N = 197) and BERT (N = 256) shapes."""
import sys; sys.path.insert(0, ".")
import torch
from nextgen_uia_b200 import ops
dev = torch.device("cuda:0")
shapes = [(256, int(a), 12) for a in sys.argv[1:]] or [(256, 197, 12), (256, 256, 12), (256, 77, 12)]
for (B, N, H) in shapes:
    D = H * 64
    qkv = torch.randn(B * N, 3 * D, device=dev).bfloat16()
    do = torch.randn(B * N, D, device=dev).bfloat16()
    o, lse = ops.attn_fwd_packed(qkv, B, N, H, 64)
    dq = ops.attn_bwd_packed(qkv, o, lse, do, B, N, H, 64)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    it = 20
    e[0].record()
    for _ in range(it):
        o, lse = ops.attn_fwd_packed(qkv, B, N, H, 64)
    e[1].record()
    for _ in range(it):
        dq = ops.attn_bwd_packed(qkv, o, lse, do, B, N, H, 64)
    e[2].record()
    torch.cuda.synchronize()
    f, b = e[0].elapsed_time(e[1]) / it, e[1].elapsed_time(e[2]) / it
    fl = 4.0 * B * H * N * N * 64
    print(f"B={B} N={N} H={H}: fwd {f*1e3:.0f} us ({fl/f/1e9:.0f} TFLOP/s)  bwd {b*1e3:.0f} us ({2.5*fl/b/1e9:.0f} TFLOP/s)")
