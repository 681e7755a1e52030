#!/usr/bin/env python
"""Benchmark of the Mona fine-tuning hot path (BASELINE.json metric: images/s, BiomedCLIP ViT-B/16 bf16).

    python bench.py --gpus N --steps K --warmup W            # our arm (N>1: launched under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU cores

One step = one pass of the hot path over one batch of synthetic input: encode_image (12 x [ViT-B/16 block
+ Mona]) + encode_text (frozen BERT-base, 77 tokens) + InfoNCE + backward + clip-grad-norm + AdamW update
(src/models/biomedclip/finetune.py:272-303 with accumulation 1).  Workload = BASELINE.json configs[1]:
batch 256 / GPU, 224x224 images, 77-token texts, bf16 compute, random-init weights of that architecture.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_IMAGE_MONA = 83.2e9  # vision fwd+bwd (frozen-weight dgrad only) + text fwd @77 tokens


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU (configs[1]: 256)")
    ap.add_argument("--method", default="mona", choices=["mona", "lora"])
    ap.add_argument("--depth", type=int, default=12)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", type=int, default=1, help="1: replay the whole step as a CUDA graph (default); 0: eager launches")
    ap.add_argument("--cpu-batch", type=int, default=8)
    return ap.parse_args()


# -------------------------------------------------------------------------------------------------
def synthetic_batch(B, seed, vocab=30522):
    """SURVEY.md §8(d) synthetic inputs: images rand[0,1) fp32, BERT ids [CLS]=2 ... [SEP]=3, no padding."""
    g = torch.Generator().manual_seed(seed)
    images = torch.rand(B, 3, 224, 224, generator=g)
    ids = torch.randint(5, vocab, (B, 77), generator=g)
    ids[:, 0] = 2
    ids[:, -1] = 3
    return images, ids


def build_model(method, depth, device=None, dtype=torch.bfloat16):
    from nextgen_uia_b200.biomedclip import BiomedCLIP, init_synthetic_
    from nextgen_uia_b200 import dp
    torch.manual_seed(1)
    model = BiomedCLIP(vision=dict(depth=depth), text=dict(layers=depth))
    init_synthetic_(model, seed=1)
    if method == "mona":
        dp.setup_mona(model, "baseline", 64)
    else:
        dp.setup_lora(model, r=8, alpha=32, dropout=0.1)
    if device is not None:
        model = model.to(device)
    model.train()
    return model.set_compute_dtype(dtype)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# -------------------------------------------------------------------------------------------------
def cpu_reference_rate(steps, warmup, batch, depth):
    """The reference algorithm (oracle port: oracle/functional.py) on the host cores, fp32, all threads:
    one training micro-step (encode_image + encode_text + InfoNCE + backward) per step."""
    from oracle import functional as OF
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = build_model("mona", depth)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    trainable = [n for n, p in model.named_parameters() if p.requires_grad]
    images, ids = synthetic_batch(batch, 1)
    cfg = dict(patch=16, depth=depth, heads=12, text_layers=depth, text_heads=12)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        OF.loss_and_grads(sd, images, ids, cfg, trainable, dtype=torch.float32)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    mean = sum(times) / len(times)
    return batch / mean, mean, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, min(args.steps, 3)), min(args.warmup, 1)
    rate, sec, cores = cpu_reference_rate(steps, warm, args.cpu_batch, args.depth)
    line = {
        "impl": "reference", "metric": "mona_finetune_images_per_sec", "value": rate, "unit": "images/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BiomedCLIP ViT-B/16 + Mona fine-tune micro-step (encode_image+encode_text+InfoNCE+backward), "
                               f"batch {args.cpu_batch} on host CPU, fp32", "depth": args.depth},
        "cpu_baseline": {"value": rate, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} timed steps of batch {args.cpu_batch} (reference modules cannot travel to the GPU box; "
                                   "oracle/functional.py is pinned to them by oracle/make_golden.py)"},
        "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def block_microbench(dev, B, iters=10, profile=False):
    """One ViT-B/16 encoder block with the Mona adapter and LoRA (r=8) on qkv/proj, forward + backward."""
    from nextgen_uia_b200.vit import Block
    from nextgen_uia_b200.adapters.mona import BaselineMona, BatchFirstMonaWrapper
    from nextgen_uia_b200.adapters.lora import LinearLoRA
    torch.manual_seed(3)
    blk = Block(768, 12)
    for p in blk.parameters():
        p.requires_grad = False
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
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if profile:
        torch.cuda.cudart().cudaProfilerStart()
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    if profile:
        torch.cuda.cudart().cudaProfilerStop()
    ms = e0.elapsed_time(e1) / iters
    tf = 6.10e9 * B / (ms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    return {"what": "ViT-B/16 block + Mona + LoRA(r=8,qkv+proj) fwd+bwd, [B,197,768] bf16, 6.10 GFLOP/image algorithmic",
            "ms": ms, "tflops": tf, "frac_of_measured_sustained_peak": tf / peak, "frac_of_nominal_2250": tf / 2250.0}


# -------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from nextgen_uia_b200 import _lib as L, ops, dp
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl ours) needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L.check(L.lib().ngu_selftest_device(), "device selftest")

    model = build_model(args.method, args.depth, dev)
    trainer = dp.Trainer(model, temperature=0.07, lr=1e-4, betas=(0.9, 0.95), weight_decay=0.01, grad_clip=1.0, accumulation_steps=1)
    B = args.batch
    images_h, ids_h = synthetic_batch(B, 1 + rank)
    images_h, ids_h = images_h.pin_memory(), ids_h.pin_memory()
    images_d, ids_d = images_h.to(dev), ids_h.to(dev)
    h2d = images_h.numel() * 4 + ids_h.numel() * 8

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, finish=None):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if finish is not None:
            finish()
        e1.record()
        sync()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    graph_note = "eager"
    if args.graph:
        try:
            trainer.capture(images_d, ids_d)
            graph_note = "cuda_graph"
        except Exception as e:  # e.g. a collective that cannot be captured on this stack: measure eagerly, say so
            trainer.graph = None
            graph_note = f"eager (capture failed: {type(e).__name__}: {str(e)[:120]})"
            torch.cuda.synchronize()

    def run_step(im, tx):
        if trainer.graph is not None:
            return trainer.replay(im, tx)
        return trainer.micro_step(im, tx)

    def step_resident():
        if trainer.graph is not None:
            trainer.replay(*trainer.static_in)
        else:
            trainer.micro_step(images_d, ids_d)

    losses = []

    # End to end through the public API: every step's batch comes from pinned HOST memory through dp.DeviceFeeder (copy
    # of batch i+1 on a side stream while batch i computes) and every step's loss is read back through dp.ScalarLog.
    def host_batches():
        while True:
            yield images_h, ids_h

    feeder = dp.DeviceFeeder(host_batches(), dev)
    log = dp.ScalarLog()

    def step_e2e():
        im, tx = next(feeder)                      # H2D of this step's inputs (154 MB), counted in feeder.h2d_bytes
        log.push(run_step(im, tx))                 # D2H of this step's loss
        losses.extend(log.pop_ready())             # host sees every loss one step late; never stalls the launch queue

    def finish_e2e():
        losses.extend(log.drain())                 # all K losses are on the host before the timed region closes

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(step_resident, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    step_e2e()
    finish_e2e()
    b0, n_l0 = feeder.h2d_bytes, len(losses)
    ms_e2e = timed(step_e2e, args.steps, finish_e2e)
    # the feeder runs one batch ahead: bytes copied inside the region / steps (== one batch per step in steady state)
    h2d = (feeder.h2d_bytes - b0) // args.steps
    assert len(losses) - n_l0 == args.steps, "every timed step's loss must have reached the host"

    value = B * world * args.steps / (ms / 1e3)
    e2e = B * world * args.steps / (ms_e2e / 1e3)

    # ---- roofline of the dominant kernel (tcgen05 GEMM): algorithmic FLOPs / CUDA-event time of every launch in one step
    roof = None
    recs = []
    orig = ops.gemm

    def timed_gemm(A, Bm, **kw):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = orig(A, Bm, **kw)
        b.record()
        k2 = kw["A2"].shape[1] if kw.get("A2") is not None else 0
        recs.append((a, b, 2.0 * A.shape[0] * Bm.shape[0] * (A.shape[1] + k2)))
        return out

    ops.gemm = timed_gemm
    n0 = L.launch_count()
    trainer.micro_step(images_d, ids_d)   # eager (so each launch can be bracketed); every rank runs it (collectives); rank 0 reports
    torch.cuda.synchronize()
    launches = (L.launch_count() - n0) * args.steps      # kernels of THIS library per step x timed steps (graph replays launch the same nodes)
    ops.gemm = orig
    if rank == 0:
        t_ms = sum(a.elapsed_time(b) for a, b, _ in recs)
        fl = sum(f for _, _, f in recs)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        ach = fl / (t_ms * 1e-3) / 1e12 if t_ms > 0 else 0.0
        roof = {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05/TMA GEMM, all launches of one step)", "achieved": ach, "peak": peak,
                "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                "traffic_note": "aggregate over shapes; per-shape DRAM bytes (QKV launch: 81 MB read + 180 MB write vs 313 MB algorithmic) in profiles/r1_gemm_qkv_ncu_full_summary.txt",
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else "fallback 1.4 PFLOP/s sustained",
                "gemm_ms_per_step": t_ms, "gemm_share_of_step": t_ms / (ms / args.steps), "gemm_launches_per_step": len(recs),
                "step_tflops_algorithmic": FLOP_PER_IMAGE_MONA * B / (ms / args.steps * 1e-3) / 1e12}

    # ---- block-level figure the north_star target is stated on: one ViT-B/16 block + Mona + LoRA(r=8, qkv+proj), fwd+bwd,
    #      on [B,197,768] bf16; algorithmic 6.10 GFLOP per image
    block = None
    if rank == 0:
        block = block_microbench(dev, B)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, sec, cores = cpu_reference_rate(3, 1, args.cpu_batch, args.depth)
        cpu = {"value": rate, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": f"3 timed micro-steps (fwd+bwd) of batch {args.cpu_batch}, fp32, oracle/functional.py, {sec:.2f} s/step"}

    if rank == 0:
        line = {
            "metric": "mona_finetune_images_per_sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"BiomedCLIP ViT-B/16 + {args.method} fine-tune with InfoNCE (BASELINE.json configs[1]): batch {B}/GPU, "
                                   "224x224, 77-token texts, 12+12 layers, fwd+bwd+clip+AdamW every step",
                       "global_batch": B * world, "parallelism": f"dp{world}", "depth": args.depth,
                       "l2": "working set per step (GBs of activations) exceeds the 126 MB L2; no explicit flush needed",
                       "launch_mode": graph_note},
            "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
            "block": block,
            "loss_last": losses[-1] if losses else None,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
