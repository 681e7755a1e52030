"""Per-op device time inside the block microbench (ViT-B/16 block + Mona + LoRA, fwd+bwd): every public function of
nextgen_uia_b200.ops is bracketed by a CUDA-event pair on the current stream; the table is the mean over the timed steps."""
import sys, os, types, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nextgen_uia_b200 import ops
from nextgen_uia_b200.vit import Block
from nextgen_uia_b200.adapters.mona import BaselineMona, BatchFirstMonaWrapper
from nextgen_uia_b200.adapters.lora import LinearLoRA

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
lora = (sys.argv[2] if len(sys.argv) > 2 else "lora") == "lora"
torch.manual_seed(3)
blk = Block(768, 12)
for p in blk.parameters():
    p.requires_grad = False
if lora:
    blk.attn.qkv = LinearLoRA(blk.attn.qkv, r=8, lora_alpha=32, dropout_rate=0.0)
    blk.attn.proj = LinearLoRA(blk.attn.proj, r=8, lora_alpha=32, dropout_rate=0.0)
mona = BatchFirstMonaWrapper(BaselineMona(768, 64))
blk, mona = blk.to(dev), mona.to(dev).eval()
x = (torch.randn(B, 197, 768, device=dev) * 0.5).bfloat16().requires_grad_(True)
g = torch.randn(B, 197, 768, device=dev).bfloat16()

def step():
    y = mona(blk(x), (14, 14))
    y.backward(g)
    x.grad = None

for _ in range(3):
    step()
torch.cuda.synchronize()

rec = []
depth = [0]
def wrap(name, fn):
    def w(*a, **k):
        if depth[0]:
            return fn(*a, **k)
        depth[0] += 1
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        try:
            r = fn(*a, **k)
        finally:
            depth[0] -= 1
        e1.record()
        shp = []
        for t in list(a) + list(k.values()):
            if isinstance(t, torch.Tensor) and t.dim() >= 2:
                shp.append("x".join(map(str, t.shape)))
        flags = ",".join(f"{kk}={vv}" for kk, vv in k.items() if isinstance(vv, (int, bool, str)) and not isinstance(vv, torch.Tensor))
        fl = 2.0 * a[0].shape[0] * a[0].shape[1] * a[1].shape[0] if name == "gemm" else 0.0
        rec.append((name, " ".join(shp[:3]), flags, e0, e1, fl))
        return r
    return w

for n in dir(ops):
    f = getattr(ops, n)
    if isinstance(f, types.FunctionType) and not n.startswith("_") and f.__module__ == ops.__name__:
        setattr(ops, n, wrap(n, f))

steps = 5
for _ in range(steps):
    torch.cuda._sleep(8_000_000)   # head start for the host, so no event pair brackets an idle GPU
    step()
torch.cuda.synchronize()
agg = collections.OrderedDict()
for name, shp, flags, e0, e1, fl in rec:
    key = (name, shp, flags)
    d = agg.setdefault(key, [0.0, 0, 0.0])
    d[0] += e0.elapsed_time(e1) * 1e3
    d[1] += 1
    d[2] = fl
tot = 0.0
for (name, shp, flags), (us, n, fl) in agg.items():
    per = us / steps
    tot += per
    extra = f"  {fl * (n / steps) / (per * 1e-6) / 1e12:7.0f} TF/s" if fl else ""
    print(f"{per:8.1f} us  n={n / steps:4.1f}  {name:18s} {shp:40s} {flags}{extra}")
print(f"total {tot:.1f} us per block step (eager, event-bracketed)")
