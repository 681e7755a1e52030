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

    # ---- RNG-stream parity of the drop-in constructors ---------------------------------------------------
    torch.set_default_dtype(torch.float32)
    from nextgen_uia_b200.adapters.mona import BaselineMona as MyMona
    from nextgen_uia_b200.adapters.lora import LinearLoRA as MyLoRA
    torch.manual_seed(3); a = rmona.BaselineMona(768, 64)
    torch.manual_seed(3); b = MyMona(768, 64)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys()) and all(torch.equal(sa[k], sb[k]) for k in sa), "Mona init stream differs"
    torch.manual_seed(5); l0 = torch.nn.Linear(768, 2304); a = rlora.LinearLoRA(l0, r=8, lora_alpha=32, dropout_rate=0.1)
    torch.manual_seed(5); l1 = torch.nn.Linear(768, 2304); b = MyLoRA(l1, r=8, lora_alpha=32, dropout_rate=0.1)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys()) and all(torch.equal(sa[k], sb[k]) for k in sa), "LoRA init stream differs"
    assert [n for n, p in a.named_parameters() if p.requires_grad] == [n for n, p in b.named_parameters() if p.requires_grad]
    print("constructor RNG streams and state-dict keys identical to the reference")


if __name__ == "__main__":
    main()
