"""torchrun --nproc-per-node 2 tools/dp_graph_check.py : the data-parallel step as CUDA-graph segments with eager NCCL calls between
them (nextgen_uia_b200/_segcap.py) against the same step launched eagerly: same loss sequence, same parameters after 3 updates
(dropout p = 0), ranks bit-identical to each other, and capturing leaves parameters / optimiser state untouched."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from nextgen_uia_b200.biomedclip import BiomedCLIP, init_synthetic_
from nextgen_uia_b200 import dp

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


def trainer():
    torch.manual_seed(1)
    model = BiomedCLIP(vision=dict(depth=2), text=dict(layers=2, vocab=1000, max_pos=128))
    init_synthetic_(model, seed=1)
    dp.setup_mona(model, "baseline", 64)
    model = model.to(dev).train().set_compute_dtype(torch.bfloat16)
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    return dp.Trainer(model, lr=1e-3, total_updates=10)


Bl = 4
g = torch.Generator().manual_seed(2 + rank)
images = torch.rand(Bl, 3, 224, 224, generator=g).to(dev)
ids = torch.randint(5, 1000, (Bl, 77), generator=g); ids[:, 0] = 2; ids[:, -1] = 3
ids = ids.to(dev)

te = trainer()
le = [float(te.micro_step(images, ids)) for _ in range(3)]
te2 = trainer()
for _ in range(3):
    te2.micro_step(images, ids)
tg = trainer()
p0 = tg.buckets.flat_param.clone()
tg.capture(images, ids)
untouched = torch.equal(tg.buckets.flat_param, p0) and tg.optimizer.updates == 0
lg = [float(tg.replay(images, ids)) for _ in range(3)]
torch.cuda.synchronize()
upd = (te.buckets.flat_param - p0).norm()
d_ee = float((te2.buckets.flat_param - te.buckets.flat_param).norm() / upd)
d_ge = float((tg.buckets.flat_param - te.buckets.flat_param).norm() / upd)
# every rank must hold the same parameters after the all-reduced updates
ref = tg.buckets.flat_param.clone()
dist.broadcast(ref, src=0)
same = torch.equal(ref, tg.buckets.flat_param)
ok = (untouched and tg.optimizer.updates == 3 and max(abs(a - b) for a, b in zip(le, lg)) < 2e-3 * abs(le[0])
      and d_ge < max(3 * d_ee, 5e-2) and same)
flag = torch.tensor([int(ok)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"world={world} segments={tg.graph.segments} eager losses {le} graph losses {lg}; param distance graph-vs-eager {d_ge:.3e} "
          f"(eager-vs-eager {d_ee:.3e}); untouched by capture {untouched}; ranks identical {same}")
    print("DP_GRAPH_OK" if int(flag.item()) else "DP_GRAPH_FAIL", flush=True)
dist.barrier()
dist.destroy_process_group()
