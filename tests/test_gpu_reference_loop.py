"""The reference's OWN training loop (src/models/biomedclip/finetune.py: prepare_model + train, unmodified) on the B200 path.

`oracle/build_ref.py` stages the script into the git-ignored oracle/_ref/.  It is imported with this repository's `src.adapters`
and `src.losses` shims in place of the reference's, and with stand-ins for what the sandbox lacks: `open_clip` (returns the
kernel-backed BiomedCLIP with synthetic weights instead of downloading a checkpoint), the dataset module, tensorboard and the
logging helpers.  Everything between — Mona injection, freezing, `encode_image` / `encode_text`, `InfoNCELoss`, `.backward()`,
`clip_grad_norm_`, `torch.optim.AdamW`, the cosine schedule, the adapter-only checkpoint — is the reference's code calling the
CUDA kernels through the module API (VERDICT r1 row b)."""
import argparse
import importlib.util
import os
import sys
import types

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, "oracle", "_ref", "finetune.py")


def _stubs(monkeypatch, n_train=6, n_val=2, batch=4):
    from nextgen_uia_b200.biomedclip import BiomedCLIP, init_synthetic_

    oc = types.ModuleType("open_clip")

    def create_model_from_pretrained(name, cache_dir=None):
        assert "BiomedCLIP" in name
        torch.manual_seed(1)
        model = BiomedCLIP(vision=dict(depth=2), text=dict(layers=2, vocab=1000, max_pos=128))
        init_synthetic_(model, seed=1, std=0.02)
        model.set_compute_dtype(torch.bfloat16)       # bf16 product path; parameters stay fp32 (the script calls model.float())
        return model, None

    def get_tokenizer(name):
        def tok(texts):
            ids = torch.zeros(len(texts), 77, dtype=torch.long)
            for i, t in enumerate(texts):
                g = torch.Generator().manual_seed(sum(ord(ch) * (j + 1) for j, ch in enumerate(t)) % (2 ** 31))
                n = 8 + len(t) % 40
                ids[i, :n] = torch.randint(5, 1000, (n,), generator=g)
                ids[i, 0] = 2
            return ids
        return tok

    oc.create_model_from_pretrained, oc.get_tokenizer = create_model_from_pretrained, get_tokenizer
    monkeypatch.setitem(sys.modules, "open_clip", oc)

    g = torch.Generator().manual_seed(7)
    scale = torch.linspace(0.2, 1.0, batch).view(-1, 1, 1, 1)
    train = [(torch.rand(batch, 3, 224, 224, generator=g) * scale, [f"ultrasound caption {b}-{i}" for i in range(batch)]) for b in range(2)]
    train = [train[i % 2] for i in range(n_train)]                         # two distinct batches, repeated: the loss must go down
    val = train[:n_val]

    ds = types.ModuleType("src.datasets")
    dsf = types.ModuleType("src.datasets.finetune")

    class DataModule:
        def __init__(self, args):
            pass

        def train_dataloader(self):
            return train

        def val_dataloader(self):
            return val

    dsf.DataModule = DataModule
    ds.finetune = dsf
    monkeypatch.setitem(sys.modules, "src.datasets", ds)
    monkeypatch.setitem(sys.modules, "src.datasets.finetune", dsf)

    ut = types.ModuleType("src.utils")
    utt = types.ModuleType("src.utils.tools")
    utt.model_summary = lambda d: "model summary (stub)"
    utt.setup_logging = lambda args, path: None
    ut.tools = utt
    monkeypatch.setitem(sys.modules, "src.utils", ut)
    monkeypatch.setitem(sys.modules, "src.utils.tools", utt)

    scalars = []
    tb = types.ModuleType("torch.utils.tensorboard")

    class SummaryWriter:
        def __init__(self, *a, **k):
            pass

        def add_scalar(self, tag, value, step):
            scalars.append((tag, float(value), int(step)))

        def close(self):
            pass

    tb.SummaryWriter = SummaryWriter
    monkeypatch.setitem(sys.modules, "torch.utils.tensorboard", tb)
    return scalars


@pytest.mark.parametrize("method", ["mona", "lora"])
def test_reference_finetune_loop_runs_on_the_kernels(method, monkeypatch, tmp_path):
    if not os.path.exists(SCRIPT):
        pytest.skip("oracle/_ref/finetune.py not staged (python oracle/build_ref.py needs /root/reference)")
    from nextgen_uia_b200 import _lib as L
    scalars = _stubs(monkeypatch)
    monkeypatch.setattr(sys, "path", list(sys.path))           # the script prepends its own project root
    spec = importlib.util.spec_from_file_location("ngu_ref_finetune", SCRIPT)
    ft = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ft)
    import src.adapters, src.losses
    assert ft.InfoNCELoss is src.losses.InfoNCELoss and ft.inject_mona_variant_to_open_clip is src.adapters.inject_mona_variant_to_open_clip

    args = argparse.Namespace(
        exp="t", method=method, tune_text_encoder=False, tune_layers="all", mona_variant="baseline", mona_bottleneck=64,
        mona_layers=None, lora_r=8, lora_alpha=32, lora_dropout=0.0, lora_layers=None, temperature=0.07, seed=1, epochs=3,
        batch_size=4, lr=2e-3, lr_min=1e-8, weight_decay=0.01, beta1_adam=0.9, beta2_adam=0.95, device="cuda:0", patience=10,
        accumulation_steps=2, grad_clip=1.0, train_snapshot_path=str(tmp_path), img_size=224, num_workers=0)
    n0 = L.launch_count()
    assert ft.train(args) == "Training Finished!"
    assert L.launch_count() - n0 > 500, "the loop must have run on this library's kernels"

    # 3 epochs x ceil(6 / 2) updates, each logged by the reference loop; the two repeated batches are being fitted
    upd = [v for tag, v, _ in scalars if tag == "train/loss_per_update"]
    assert len(upd) == 9 and all(torch.isfinite(torch.tensor(upd)))
    assert upd[-1] < upd[0] - 1e-3, upd
    lrs = [v for tag, v, _ in scalars if tag == "train/lr"]
    assert lrs[0] > lrs[-1] >= 0.0                                  # CosineAnnealingLR stepped once per update
    # adapter-only checkpoint with the reference's key filter
    sd = torch.load(os.path.join(str(tmp_path), "best_model.pth"))
    assert len(sd) > 0 and all(method in k.lower() for k in sd)
    assert all(torch.isfinite(v).all() for v in sd.values())
