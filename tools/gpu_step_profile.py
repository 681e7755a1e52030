"""In-situ per-kernel GPU time of the benchmark step (torch.profiler / CUPTI: concurrent, real clocks, warm caches —
unlike the serialised cold-cache ncu launch list).  usage: python tools/gpu_step_profile.py [mona|lora] [steps]"""
import collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from nextgen_uia_b200 import dp
from torch.profiler import profile, ProfilerActivity

method = sys.argv[1] if len(sys.argv) > 1 else "mona"      # mona | lora | cfg4
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
if method == "cfg4":
    model = bench.build_clip_model(4, dev)
    tr = dp.Trainer(model)
    im, ids = torch.rand(64, 3, 336, 336), bench.clip_tokens(64, 1)
else:
    model = bench.build_model(method, 12, dev)
    tr = dp.Trainer(model)
    im, ids = bench.synthetic_batch(256, 1)
im, ids = im.to(dev), ids.to(dev)
for _ in range(4):
    tr.micro_step(im, ids)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        tr.micro_step(im, ids)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").split("(")[0][:90]
        agg[name][0] += 1
        agg[name][1] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
tot = sum(v[1] for v in agg.values())
print(f"total kernel time {tot / steps:.0f} us/step over {sum(v[0] for v in agg.values()) // steps} launches/step")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{v[1] / steps:10.0f} us {100 * v[1] / tot:5.1f}% n={v[0] // steps:4d} avg {v[1] / v[0]:8.1f} us  {k}")
