"""Scratch: tcgen05 attention vs SIMT vs torch SDPA (fp64 CPU), plus timing at cfg2 shape."""
import sys
sys.path.insert(0, ".")
import torch, torch.nn.functional as F
from nextgen_uia_b200 import ops
dev = torch.device("cuda:0")

def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max())

for (B, N, H) in [(2, 197, 12), (3, 77, 12), (1, 128, 2), (2, 256, 3), (5, 130, 1), (3, 64, 2), (2, 16, 1), (4, 65, 3), (2, 192, 2), (3, 193, 2), (30, 197, 12), (40, 77, 12), (70, 208, 5)]:
    torch.manual_seed(0)
    dh, D = 64, H * 64
    qkv = torch.randn(B * N, 3 * D).to(dev, torch.bfloat16)
    do = torch.randn(B * N, D).to(dev, torch.bfloat16)
    t = qkv.double().cpu().requires_grad_(True)
    q, k, v = t.view(B, N, 3, H, dh).permute(2, 0, 3, 1, 4)
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * N, D)
    (dref,) = torch.autograd.grad((ref * do.double().cpu()).sum(), t)
    try:
        o, lse = ops.attn_fwd_packed(qkv, B, N, H, dh, impl=0)
        torch.cuda.synchronize()
        o1, lse1 = ops.attn_fwd_packed(qkv, B, N, H, dh, impl=1)
        print(f"B={B} N={N} H={H}: fwd tc relerr {rel(o, ref):.3e} (simt {rel(o1, ref):.3e}) lse err {rel(lse, lse1):.3e}", flush=True)
        dq = ops.attn_bwd_packed(qkv, o1, lse1, do, B, N, H, dh, impl=0)
        torch.cuda.synchronize()
        dq1 = ops.attn_bwd_packed(qkv, o1, lse1, do, B, N, H, dh, impl=1)
        for nm, sl in (("dq", slice(0, D)), ("dk", slice(D, 2 * D)), ("dv", slice(2 * D, 3 * D))):
            print(f"    bwd {nm}: tc relerr {rel(dq[:, sl], dref[:, sl]):.3e} (simt {rel(dq1[:, sl], dref[:, sl]):.3e})", flush=True)
    except Exception as e:
        print("EXC", e, flush=True)
        break

B, N, H, dh = 256, 197, 12, 64
D = H * dh
qkv = torch.randn(B * N, 3 * D).to(dev, torch.bfloat16)
do = torch.randn(B * N, D).to(dev, torch.bfloat16)
o, lse = ops.attn_fwd_packed(qkv, B, N, H, dh)
def tm(fn, it=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it
f = tm(lambda: ops.attn_fwd_packed(qkv, B, N, H, dh))
bw = tm(lambda: ops.attn_bwd_packed(qkv, o, lse, do, B, N, H, dh))
fl = 4.0 * B * H * N * N * dh
print(f"cfg2 attention: fwd {f*1e3:.0f} us ({fl/f/1e9:.0f} TF/s)  bwd {bw*1e3:.0f} us ({2.5*fl/bw/1e9:.0f} TF/s)")
q4 = qkv.view(B, N, 3, H, dh).permute(2, 0, 3, 1, 4).contiguous()
qq, kk, vv = q4[0].requires_grad_(True), q4[1].requires_grad_(True), q4[2].requires_grad_(True)
def sd():
    return F.scaled_dot_product_attention(qq, kk, vv)
ft = tm(sd)
oo = sd(); g = torch.randn_like(oo)
bt = tm(lambda: torch.autograd.grad(oo, [qq, kk, vv], g, retain_graph=True))
print(f"torch SDPA (flash) fwd {ft*1e3:.0f} us bwd {bt*1e3:.0f} us")
