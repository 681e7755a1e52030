import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nextgen_uia_b200.biomedclip import BiomedCLIP, init_synthetic_
from nextgen_uia_b200 import dp
from oracle import functional as OF
dev = torch.device("cuda:0")
torch.manual_seed(1)
model = BiomedCLIP(vision=dict(depth=2), text=dict(layers=2, vocab=1000, max_pos=128))
init_synthetic_(model, seed=1)
dp.setup_mona(model, "baseline", 64)
sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
trainable = [n for n, p in model.named_parameters() if p.requires_grad]
g = torch.Generator().manual_seed(2)
images = torch.rand(4, 3, 224, 224, generator=g) * torch.linspace(0.2, 1.0, 4).view(-1, 1, 1, 1)
ids = torch.randint(5, 1000, (4, 77), generator=g); ids[:, 0] = 2; ids[:, -1] = 3
cfg = dict(patch=16, depth=2, heads=12, text_layers=2, text_heads=12)
model = model.to(dev).eval().set_compute_dtype(torch.float32)
tr = dp.Trainer(model, grad_clip=0.0, lr=0.0)
print("sink marked:", sum(1 for p in model.parameters() if getattr(p, "_ngu_sink", None) is not None), "of", len(tr.params))
fi = model.encode_image(images.to(dev)); ft = model.encode_text(ids.to(dev))
loss = tr.criterion(fi, ft)
loss.backward()
torch.cuda.synchronize()
lo, _, _, _, go = OF.loss_and_grads(sd, images, ids, cfg, trainable)
for n, p in list(model.named_parameters()):
    if p.requires_grad and ("blocks.1.mona" in n):
        print(n.split("clip_mona.")[-1], float(p.grad.norm()), float(go[n].norm()))
