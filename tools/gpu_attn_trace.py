"""Timeline of the attention-backward roles on block 0 (needs the library built with -DNGU_ATTN_TRACE)."""
import sys, ctypes
sys.path.insert(0, ".")
import torch, numpy as np
from nextgen_uia_b200 import ops, _lib as L
lib = L.lib()
dev = torch.device("cuda:0")
B, N, H, dh = 256, 197, 12, 64
D = H * dh
qkv = torch.randn(B * N, 3 * D).to(dev, torch.bfloat16)
do = torch.randn(B * N, D).to(dev, torch.bfloat16)
o, lse = ops.attn_fwd_packed(qkv, B, N, H, dh)
dq = ops.attn_bwd_packed(qkv, o, lse, do, B, N, H, dh)
torch.cuda.synchronize()
lib.ngu_debug_attn_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
lib.ngu_debug_attn_trace(None, 1)
if "fwd" in sys.argv:
    ops.attn_fwd_packed(qkv, B, N, H, dh)
else:
    dq = ops.attn_bwd_packed(qkv, o, lse, do, B, N, H, dh)
torch.cuda.synchronize()
buf = np.zeros(32 * 256 * 3, dtype=np.uint64)
lib.ngu_debug_attn_trace(buf.ctypes.data, 0)
ev = buf.reshape(-1, 3).astype(np.int64)
ev = ev[ev[:, 2] > 0]
n = len(ev)
t0 = ev[:, 2].min()
names3 = {40: "sm.waitS", 42: "sm.Sseen", 44: "sm.maxdone", 46: "sm.barrier", 48: "sm.Parrive", 50: "ep.waitO", 51: "ep.Oseen", 52: "ep.drained", 53: "ep.staged", 60: "M.S", 62: "M.PV", 61: "M.Pseen", 54: "C.stored", 55: "C.loads", 59: "C.Sready", 49: "sm15.Parrive", 47: "sm12.Parrive"}
names = {40: "g0.waitS", 41: "g1.waitS", 42: "g0.Sseen", 43: "g1.Sseen", 44: "g0.pass1", 45: "g1.pass1", 46: "g0.pass2", 47: "g1.pass2", 48: "g0.Parr", 49: "g1.Parr", 50: "g0.Oseen", 51: "g1.Oseen", 52: "g0.drained", 53: "g1.drained", 60: "M.S(g0)", 61: "M.S(g1)", 62: "M.PV(g0)", 63: "M.PV(g1)", 10: "S.issue.begin", 11: "S.issued", 20: "M.waitP", 21: "M.Pseen", 22: "M.issued", 30: "C.waitS", 31: "C.Sseen", 32: "C.Parrive", 33: "C.drained"}
if "v3" in sys.argv:
    names = names3
ev = ev[np.argsort(ev[:, 2])]
nums = [int(a) for a in sys.argv[1:] if a.isdigit()]
lo, hi = nums[0] if len(nums) > 0 else 16, nums[1] if len(nums) > 1 else 34
for c, a, t in ev:
    if lo <= a < hi:
        print(f"{t - t0:8d}  step {a:3d}  {names.get(int(c), c)}")
print("events", n, "span cycles", int(ev[:, 2].max() - t0))
