"""Run the bench workload and bracket ONE steady-state step with cudaProfilerStart/Stop
(use with: ncu --profile-from-start off ...)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from nextgen_uia_b200 import dp
method = sys.argv[1] if len(sys.argv) > 1 else "mona"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda:0")
model = bench.build_model(method, 12, dev)
tr = dp.Trainer(model)
im, ids = bench.synthetic_batch(B, 1)
im, ids = im.to(dev), ids.to(dev)
for _ in range(3):
    tr.micro_step(im, ids)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
tr.micro_step(im, ids)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
