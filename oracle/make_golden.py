"""Pin the oracle against the reference's own modules and write tests/golden/*.pt.

Runs ONLY in the build container (needs /root/reference).  It imports the UNMODIFIED reference files
src/adapters/mona.py, src/adapters/lora.py, src/losses/losses.py by path (the reference package
__init__ is broken as shipped), executes them on seeded CPU inputs and
  1. asserts the functional oracle (oracle/functional.py) reproduces them (fp64, 1e-10), and
  2. stores small golden input/param/output/grad vectors (fp32) for the tests that run without
     /root/reference (GPU box, CI).
Usage:  python oracle/make_golden.py
"""
import importlib.util
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
from oracle import functional as OF  # noqa: E402


def load_ref(rel, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    rmona = load_ref("src/adapters/mona.py", "ref_mona")
    rlora = load_ref("src/adapters/lora.py", "ref_lora")
    rloss = load_ref("src/losses/losses.py", "ref_losses")
    torch.set_default_dtype(torch.float64)

    # ---- Mona (batch-first wrapper, CLS + 14x14 grid, and the no-CLS 4x4 path) -------------------
    for tag, (B, hw, D, has_cls) in {"mona_cls": (2, (14, 14), 256, True), "mona_nocls": (3, (4, 4), 256, False)}.items():
        torch.manual_seed(7)
        m = rmona.BatchFirstMonaWrapper(rmona.BaselineMona(D, 64)).double().eval()
        with torch.no_grad():
            m.clip_mona.gamma.copy_(torch.randn(D) * 0.5)  # gamma init 1e-6 would hide the LN branch
            m.clip_mona.gammax.copy_(1 + 0.1 * torch.randn(D))
            m.clip_mona.norm.weight.copy_(1 + 0.1 * torch.randn(D))
            m.clip_mona.norm.bias.copy_(0.1 * torch.randn(D))
        N = hw[0] * hw[1] + (1 if has_cls else 0)
        x = torch.randn(B, N, D, requires_grad=True)
        gy = torch.randn(B, N, D)
        y = m(x, hw if has_cls else None)
        params = list(m.named_parameters())
        grads = torch.autograd.grad((y * gy).sum(), [x] + [p for _, p in params])
        sd = {f"clip_mona.{k}" if not k.startswith("clip_mona.") else k: v.detach() for k, v in m.state_dict().items()}
        # oracle check
        p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        xo = x.detach().clone().requires_grad_(True)
        yo = OF.mona(xo, p, "clip_mona.", hw, has_cls)
        go = torch.autograd.grad((yo * gy).sum(), [xo] + [p[k] for k, _ in params])
        assert torch.allclose(yo, y, atol=1e-10, rtol=1e-10), "oracle mona fwd != reference"
        for a, b in zip(go, grads):
            assert torch.allclose(a, b, atol=1e-9, rtol=1e-9), "oracle mona grad != reference"
        torch.save({"state": {k: v.float() for k, v in sd.items()}, "x": x.detach().float(), "gy": gy.float(), "hw": hw,
                    "has_cls": has_cls, "y": y.detach().float(), "dx": grads[0].float(),
                    "grads": {k: g.float() for (k, _), g in zip(params, grads[1:])}}, os.path.join(out_dir, f"{tag}.pt"))
        print(tag, "ok; reference == oracle (fp64), golden written")

    # ---- Mona variants (noise-aware / frequency-enhanced / hybrid), 6x6 grid + CLS ---------------------------
    for tag, cls in {"mona_noise": rmona.NoiseAwareMona, "mona_freq": rmona.FreqEnhancedMona, "mona_hybrid": rmona.HybridNoiseFreqMona}.items():
        torch.manual_seed(23)
        D, hw, B = 256, (6, 6), 3
        m = rmona.BatchFirstMonaWrapper(cls(D, 64)).double().eval()
        with torch.no_grad():
            m.clip_mona.gamma.copy_(torch.randn(D) * 0.5)
            if hasattr(m.clip_mona.adapter_conv, "freq_filter"):
                m.clip_mona.adapter_conv.freq_filter.copy_(1 + 0.3 * torch.randn(64))
            if hasattr(m.clip_mona.adapter_conv, "noise_estimator"):
                for prm in m.clip_mona.adapter_conv.noise_estimator.parameters():
                    prm.mul_(3.0)  # spread the softmax so the branch weights are not ~1/3
        N = hw[0] * hw[1] + 1
        x = torch.randn(B, N, D, requires_grad=True)
        gy = torch.randn(B, N, D)
        y = m(x, hw)
        params = list(m.named_parameters())
        grads = torch.autograd.grad((y * gy).sum(), [x] + [p for _, p in params])
        sd = {k: v.detach() for k, v in m.state_dict().items()}
        p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        xo = x.detach().clone().requires_grad_(True)
        yo = OF.mona(xo, p, "clip_mona.", hw, True)
        go = torch.autograd.grad((yo * gy).sum(), [xo] + [p[k] for k, _ in params])
        assert torch.allclose(yo, y, atol=1e-10, rtol=1e-10), tag
        for a, b in zip(go, grads):
            assert torch.allclose(a, b, atol=1e-9, rtol=1e-9), tag
        torch.save({"state": {k: v.float() for k, v in sd.items()}, "x": x.detach().float(), "gy": gy.float(), "hw": hw,
                    "has_cls": True, "y": y.detach().float(), "dx": grads[0].float(),
                    "grads": {k: g_.float() for (k, _), g_ in zip(params, grads[1:])}}, os.path.join(out_dir, f"{tag}.pt"))
        print(tag, "ok; reference == oracle (fp64), golden written")

    # ---- LinearLoRA ----------------------------------------------------------------------------------
    torch.manual_seed(11)
    lin = torch.nn.Linear(256, 384).double()
    ll = rlora.LinearLoRA(lin, r=8, lora_alpha=32, dropout_rate=0.1).double().eval()
    with torch.no_grad():
        ll.w_lora_B.copy_(torch.randn(384, 8) * 0.02)  # B = 0 at init would leave the branch untested
    x = torch.randn(5, 7, 256, requires_grad=True)
    gy = torch.randn(5, 7, 384)
    y = ll(x)
    names = ["w_lora_A", "w_lora_B", "bias"]
    grads = torch.autograd.grad((y * gy).sum(), [x] + [getattr(ll, n) for n in names])
    sd = {k: v.detach() for k, v in ll.state_dict().items()}
    p = {f"l.{k}": v.clone().requires_grad_(True) for k, v in sd.items()}
    xo = x.detach().clone().requires_grad_(True)
    yo = OF.lora_linear(xo, p, "l.", 8, 32)
    go = torch.autograd.grad((yo * gy).sum(), [xo] + [p[f"l.{n}"] for n in names])
    assert torch.allclose(yo, y, atol=1e-10, rtol=1e-10)
    for a, b in zip(go, grads):
        assert torch.allclose(a, b, atol=1e-9, rtol=1e-9)
    assert abs(ll.scaling - 32 / 8 ** 0.5) < 1e-12
    torch.save({"state": {k: v.float() for k, v in sd.items()}, "x": x.detach().float(), "gy": gy.float(), "y": y.detach().float(),
                "dx": grads[0].float(), "grads": {n: g.float() for n, g in zip(names, grads[1:])}, "r": 8, "alpha": 32},
               os.path.join(out_dir, "lora_linear.pt"))
    print("lora_linear ok; reference == oracle (fp64), golden written")

    # ---- InfoNCE ---------------------------------------------------------------------------------------
    torch.manual_seed(13)
    for tag, B in {"infonce_b8": 8, "infonce_b37": 37}.items():
        I = torch.randn(B, 512, requires_grad=True)
        T = torch.randn(B, 512, requires_grad=True)
        crit = rloss.InfoNCELoss(temperature=0.07)
        loss = crit(I, T)
        gI, gT = torch.autograd.grad(loss, [I, T])
        Io, To = I.detach().clone().requires_grad_(True), T.detach().clone().requires_grad_(True)
        lo, _ = OF.info_nce(Io, To, 0.07)
        gIo, gTo = torch.autograd.grad(lo, [Io, To])
        assert torch.allclose(lo, loss, atol=1e-12) and torch.allclose(gIo, gI, atol=1e-12) and torch.allclose(gTo, gT, atol=1e-12)
        torch.save({"I": I.detach().float(), "T": T.detach().float(), "loss": loss.detach().float(), "dI": gI.float(), "dT": gT.float(),
                    "temperature": 0.07}, os.path.join(out_dir, f"{tag}.pt"))
        print(tag, "ok; reference == oracle (fp64), golden written")

    # ---- vendored OpenAI CLIP (ViT tower with Mona, and with LoRA) ------------------------------------------
    rclip = load_ref("src/third_party/openai_clip/model.py", "ref_clip_model")
    cfg = dict(patch=16, depth=1, heads=4, text_layers=1, text_heads=1)
    for tag in ("clip_mona", "clip_lora"):
        torch.manual_seed(17)
        # the vendored LayerNorm casts to fp32 (model.py:168), so the reference tower runs in fp32; the oracle runs in fp64
        m = rclip.CLIP(64, 32, 1, 256, 16, 8, 50, 64, 1, 1).float()
        with torch.no_grad():  # bf16-representable weights keep the fixture small (stored as bf16)
            for prm in m.parameters():
                prm.copy_(prm.bfloat16().float())
        for prm in m.parameters():
            prm.requires_grad = False
        if tag == "clip_mona":
            rmona.inject_mona_variant_to_clip(m, variant="baseline", bottleneck_dim=64)
            key = "mona"
        else:
            rlora.inject_lora_to_clip(m, lora_r=8, lora_alpha=32, lora_dropout=0.1)
            key = "lora"
        m = m.float().eval()
        with torch.no_grad():
            for n, prm in m.named_parameters():
                if n.endswith("gamma"):
                    prm.copy_(torch.randn(prm.shape) * 0.3)
                if n.endswith("w_lora_B"):
                    prm.copy_(torch.randn(prm.shape) * 0.05)
                if key in n:
                    prm.copy_(prm.bfloat16().float())
        trainable = [n for n, _ in m.named_parameters() if key in n.lower()]
        for n, prm in m.named_parameters():
            prm.requires_grad = n in trainable
        images = torch.rand(3, 3, 32, 32).float()
        text = torch.randint(1, 48, (3, 8)); text[:, -1] = 49
        gi = torch.randn(3, 64).float()
        fi = m.encode_image(images)
        with torch.no_grad():
            ft = m.encode_text(text)
        grads = torch.autograd.grad((fi * gi).sum(), [dict(m.named_parameters())[n] for n in trainable])
        sd = {k: v.detach() for k, v in m.state_dict().items()}
        p = {k: (v.double().clone().requires_grad_(k in trainable) if v.is_floating_point() else v) for k, v in sd.items()}
        c2 = dict(cfg, lora=(8, 32)) if tag == "clip_lora" else cfg
        fo = OF.clip_encode_image(p, images.double(), c2)
        to = OF.clip_encode_text(p, text, c2)
        go = torch.autograd.grad((fo * gi.double()).sum(), [p[n] for n in trainable])
        rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
        assert rel(fo, fi) < 1e-5 and rel(to, ft) < 1e-5, (tag, rel(fo, fi), rel(to, ft))
        for n_, a, b in zip(trainable, go, grads):
            assert rel(a, b) < 1e-4, (tag, n_, rel(a, b))
        torch.save({"state": {k: (v.bfloat16() if v.is_floating_point() else v) for k, v in sd.items()}, "images": images.float(), "text": text,
                    "gi": gi.float(), "fi": fi.detach().float(), "ft": ft.float(), "trainable": trainable,
                    "grads": {n: g_.float() for n, g_ in zip(trainable, grads)}, "cfg": c2}, os.path.join(out_dir, f"{tag}.pt"))
        print(tag, "ok; vendored reference CLIP == oracle (fp64), golden written")

    # ---- RNG-stream parity of the drop-in constructors ---------------------------------------------------
    torch.set_default_dtype(torch.float32)
    from nextgen_uia_b200.adapters.mona import BaselineMona as MyMona
    from nextgen_uia_b200.adapters.lora import LinearLoRA as MyLoRA
    torch.manual_seed(3); a = rmona.BaselineMona(768, 64)
    torch.manual_seed(3); b = MyMona(768, 64)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys()) and all(torch.equal(sa[k], sb[k]) for k in sa), "Mona init stream differs"
    from nextgen_uia_b200.adapters import mona as my_mona
    for nm in ("NoiseAwareMona", "FreqEnhancedMona", "HybridNoiseFreqMona"):
        torch.manual_seed(4); a = getattr(rmona, nm)(768, 64)
        torch.manual_seed(4); b = getattr(my_mona, nm)(768, 64)
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa.keys()) == list(sb.keys()) and all(torch.equal(sa[k], sb[k]) for k in sa), nm + " init stream differs"
    torch.manual_seed(5); l0 = torch.nn.Linear(768, 2304); a = rlora.LinearLoRA(l0, r=8, lora_alpha=32, dropout_rate=0.1)
    torch.manual_seed(5); l1 = torch.nn.Linear(768, 2304); b = MyLoRA(l1, r=8, lora_alpha=32, dropout_rate=0.1)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys()) and all(torch.equal(sa[k], sb[k]) for k in sa), "LoRA init stream differs"
    assert [n for n, p in a.named_parameters() if p.requires_grad] == [n for n, p in b.named_parameters() if p.requires_grad]
    print("constructor RNG streams and state-dict keys identical to the reference")


if __name__ == "__main__":
    main()
