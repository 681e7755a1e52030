"""Golden vectors at the BENCHMARKED depth: BASELINE.json configs[0] and configs[1] shapes.

Runs the CPU oracle (oracle/functional.py, pinned to the reference modules by oracle/make_golden.py) on exactly the
model `bench.build_model("mona", 12)` builds and the inputs `bench.synthetic_batch` draws:

  tests/golden/cfg1_b8_d12.pt    batch 8, 12+12 layers, fp64: image/text features, loss, per-layer residual-stream taps
                                 (first 2 images, 5 tokens), the FULL adapter gradients of layers 0, 5, 11 and the
                                 L2 norm of every adapter gradient tensor.
  tests/golden/cfg2_b256_d12.pt  batch 256, 12+12 layers, fp32 oracle (fp64 would take ~10 min here): features, loss,
                                 per-layer taps (first 2 images, 5 tokens).

The GPU box has no /root/reference and the oracle at batch 256 costs ~12 TFLOP of CPU work, so these are committed
fixtures; weights are NOT stored (they are re-created bit-identically from the seed; a checksum guards drift).
Usage:  python oracle/make_golden_cfg.py [--only cfg1|cfg2]
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import functional as OF  # noqa: E402

TAP_IMGS, TAP_TOKS = 2, (0, 1, 57, 100, 196)
FULL_GRAD_LAYERS = (0, 5, 11)


def weight_checksum(sd):
    """Order-independent fingerprint of the synthetic weights (fp64 sums of a few tensors)."""
    keys = ["visual.trunk.blocks.0.attn.qkv.weight", "visual.trunk.blocks.11.mlp.fc2.weight", "visual.head.proj.weight",
            "text.transformer.encoder.layer.11.output.dense.weight", "text.proj.2.weight",
            "visual.trunk.blocks.11.mona.clip_mona.project2.weight", "visual.trunk.blocks.0.mona.clip_mona.adapter_conv.conv3.weight"]
    return torch.tensor([float(sd[k].double().abs().sum()) for k in keys], dtype=torch.float64)


def build(depth=12):
    import bench
    model = bench.build_model("mona", depth).eval()
    # gamma is initialised to 1e-6 (mona.py:112): that hides the LayerNorm branch of the adapter completely, so the
    # parity models use an O(0.2) gamma (deterministic), like the depth-2 tests do
    g = torch.Generator().manual_seed(11)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("gamma"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.2)
    return model


def taps_small(taps):
    return torch.stack([t[:TAP_IMGS][:, list(TAP_TOKS), :].float() for t in taps], 0)   # [depth, imgs, toks, D]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    import bench
    torch.set_num_threads(os.cpu_count() or 1)
    model = build(12)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    trainable = [n for n, p in model.named_parameters() if p.requires_grad]
    cfg = dict(patch=16, depth=12, heads=12, text_layers=12, text_heads=12)
    out_dir = os.path.join(ROOT, "tests", "golden")
    ck = weight_checksum(sd)

    if args.only in ("", "cfg1"):
        t0 = time.time()
        images, ids = bench.synthetic_batch(8, 1)
        p = {k: v.double().clone() if v.is_floating_point() else v.clone() for k, v in sd.items()}
        for k in trainable:
            p[k].requires_grad_(True)
        taps = []
        fi = OF.encode_image(p, images.double(), cfg, taps=taps)
        with torch.no_grad():
            ft = OF.encode_text(p, ids, cfg)
        loss, logits = OF.info_nce(fi, ft, 0.07)
        grads = torch.autograd.grad(loss, [p[k] for k in trainable], retain_graph=True)
        # a well-conditioned cotangent for the bf16 tower-backward check (see tests/test_gpu_modules.py)
        G = torch.randn(fi.shape, generator=torch.Generator().manual_seed(3), dtype=torch.float64)
        gradsG = torch.autograd.grad((fi * G).sum(), [p[k] for k in trainable])
        full = {k: g.float() for k, g in zip(trainable, grads) if any(f".blocks.{i}." in k for i in FULL_GRAD_LAYERS)}
        fullG = {k: g.float() for k, g in zip(trainable, gradsG) if any(f".blocks.{i}." in k for i in FULL_GRAD_LAYERS)}
        torch.save({"fi": fi.detach().float(), "ft": ft.float(), "loss": float(loss), "logits": logits.detach().float(),
                    "taps": taps_small([t.detach() for t in taps]), "grads": full, "gradsG": fullG, "G": G.float(),
                    "grad_norms": {k: float(g.norm()) for k, g in zip(trainable, grads)},
                    "gradG_norms": {k: float(g.norm()) for k, g in zip(trainable, gradsG)},
                    "checksum": ck, "oracle_dtype": "float64"}, os.path.join(out_dir, "cfg1_b8_d12.pt"))
        print(f"cfg1_b8_d12: loss {float(loss):.6f}  ({time.time() - t0:.0f} s)")

    if args.only in ("", "cfg2"):
        t0 = time.time()
        images, ids = bench.synthetic_batch(256, 1)
        with torch.no_grad():
            p = {k: v.float().clone() if v.is_floating_point() else v.clone() for k, v in sd.items()}
            taps = []
            fis, fts = [], []
            for s in range(0, 256, 32):       # chunks keep the attention temporaries small; the math is per image
                tp = []
                fis.append(OF.encode_image(p, images[s:s + 32], cfg, taps=tp))
                fts.append(OF.encode_text(p, ids[s:s + 32], cfg))
                if s == 0:
                    taps = tp
            fi, ft = torch.cat(fis), torch.cat(fts)
            loss, _ = OF.info_nce(fi.double(), ft.double(), 0.07)
        torch.save({"fi": fi, "ft": ft, "loss": float(loss), "taps": taps_small(taps), "checksum": ck, "oracle_dtype": "float32"},
                   os.path.join(out_dir, "cfg2_b256_d12.pt"))
        print(f"cfg2_b256_d12: loss {float(loss):.6f}  ({time.time() - t0:.0f} s)")


if __name__ == "__main__":
    main()
