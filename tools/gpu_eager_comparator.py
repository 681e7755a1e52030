"""On-box GPU comparator: the oracle restatement of the reference step -- plain PyTorch ops, i.e. cuBLASLt
GEMMs, native LayerNorm / depthwise conv / softmax kernels -- run on the B200 under bf16 autocast, timed beside the
hand-written kernels.  A tool, not part of the product or of bench.py: it is the realistic bar an unmodified PyTorch port
of the reference would set on this box.  Usage: python tools/gpu_eager_comparator.py [batch] [steps]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from oracle import functional as OF

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda:0")
model = bench.build_model("mona", 12, dev)
sd = {k: v.detach() for k, v in model.state_dict().items()}
trainable = [n for n, p in model.named_parameters() if p.requires_grad]
cfg = dict(patch=16, depth=12, heads=12, text_layers=12, text_heads=12)
images, ids = bench.synthetic_batch(B, 1)
images, ids = images.to(dev), ids.to(dev)
p = {k: (v.clone().requires_grad_(k in trainable) if v.is_floating_point() else v) for k, v in sd.items()}
opt = torch.optim.AdamW([p[k] for k in trainable], lr=1e-4)


def step():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss, fi, ft, logits = OF.training_loss(p, images, ids, cfg)
    loss.backward()
    torch.nn.utils.clip_grad_norm_([p[k] for k in trainable], 1.0)
    opt.step()
    opt.zero_grad(set_to_none=True)
    return loss


for _ in range(2):
    l = step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(steps):
    l = step()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / steps
print(f"torch eager (autocast bf16) oracle step: batch {B}: {ms:.1f} ms/step = {B / ms * 1e3:.0f} images/s  (loss {float(l):.4f}, "
      f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB)")
