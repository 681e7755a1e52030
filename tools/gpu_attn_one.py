"""Launch the tcgen05 attention forward and backward once each at the ViT-B/16 shape (ncu target)."""
import sys
sys.path.insert(0, ".")
import torch
from nextgen_uia_b200 import ops
dev = torch.device("cuda:0")
B, N, H, dh = int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 197, 12, 64
D = H * dh
qkv = torch.randn(B * N, 3 * D).to(dev, torch.bfloat16)
do = torch.randn(B * N, D).to(dev, torch.bfloat16)
for _ in range(2):
    o, lse = ops.attn_fwd_packed(qkv, B, N, H, dh)
    dq = ops.attn_bwd_packed(qkv, o, lse, do, B, N, H, dh)
torch.cuda.synchronize()
print("done")
