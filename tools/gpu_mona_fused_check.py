"""Fused Mona path (csrc/mona_fused.cu) vs the unfused kernels and the fp64 oracle, plus timing at the bench shape."""
import os, sys
sys.path.insert(0, ".")
import torch
from nextgen_uia_b200.adapters.mona import BaselineMona, FreqEnhancedMona, BatchFirstMonaWrapper
from oracle import functional as OF

dev = torch.device("cuda:0")
rel = lambda a, b: float((a.double().cpu() - b.double().cpu()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def run(m, x, gy, hw, fused):
    os.environ["NGU_MONA_FUSED"] = "1" if fused else "0"
    for p in m.parameters():
        p.grad = None
    xg = x.clone().requires_grad_(True)
    y = m(xg, hw)
    (y.float() * gy.float()).sum().backward()
    return y.detach(), xg.grad.detach(), {n: p.grad.detach().clone() for n, p in m.named_parameters()}


def check(cls, B, grid, D, has_cls=True, seed=0):
    torch.manual_seed(seed)
    m = BatchFirstMonaWrapper(cls(D, 64))
    with torch.no_grad():
        m.clip_mona.gamma.copy_(torch.randn(D) * 0.3)
        m.clip_mona.gammax.copy_(1 + 0.1 * torch.randn(D))
        m.clip_mona.norm.weight.copy_(1 + 0.1 * torch.randn(D))
        m.clip_mona.norm.bias.copy_(0.1 * torch.randn(D))
        if hasattr(m.clip_mona.adapter_conv, "freq_filter"):
            m.clip_mona.adapter_conv.freq_filter.copy_(1 + 0.2 * torch.randn(64))
    sd = {k: v.detach().double() for k, v in m.state_dict().items()}
    N = grid * grid + (1 if has_cls else 0)
    x = (torch.randn(B, N, D) * 0.7 + 0.1).bfloat16()
    gy = torch.randn(B, N, D).bfloat16()
    m = m.to(dev).eval()
    hw = (grid, grid) if has_cls else None
    yf, dxf, gf = run(m, x.to(dev), gy.to(dev), hw, True)
    yu, dxu, gu = run(m, x.to(dev), gy.to(dev), hw, False)
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xo = x.double().requires_grad_(True)
    yo = OF.mona(xo, p, "clip_mona.", (grid, grid), has_cls)
    names = [n for n, _ in m.named_parameters()]
    go = torch.autograd.grad((yo * gy.double()).sum(), [xo] + [p[n] for n in names])
    print(f"{cls.__name__} B={B} grid={grid} D={D} cls={has_cls}:  y fused {rel(yf, yo):.2e} unfused {rel(yu, yo):.2e} | "
          f"dx fused {rel(dxf, go[0]):.2e} unfused {rel(dxu, go[0]):.2e}")
    worst = 0.0
    for n, gref in zip(names, go[1:]):
        ef, eu = rel(gf[n], gref), rel(gu[n], gref)
        worst = max(worst, ef)
        flag = "  <<<<" if ef > 2e-2 and ef > 2 * eu else ""
        print(f"    {n:42s} fused {ef:.2e}  unfused {eu:.2e}{flag}")
    return worst


def timing(B=256):
    torch.manual_seed(0)
    m = BatchFirstMonaWrapper(BaselineMona(768, 64)).to(dev).train()
    x = (torch.randn(B, 197, 768, device=dev) * 0.5).bfloat16().requires_grad_(True)
    g = torch.randn(B, 197, 768, device=dev).bfloat16()
    for fused in (True, False):
        os.environ["NGU_MONA_FUSED"] = "1" if fused else "0"
        def step():
            y = m(x, (14, 14)); y.backward(g); x.grad = None
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            step()
        e1.record(); torch.cuda.synchronize()
        print(f"Mona fwd+bwd B={B} train mode, fused={fused}: {e0.elapsed_time(e1) / 10 * 1e3:.0f} us")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "time":
        timing()
        sys.exit(0)
    check(BaselineMona, 2, 14, 256)
    check(BaselineMona, 3, 4, 256, has_cls=False)
    check(FreqEnhancedMona, 2, 6, 256)
    check(BaselineMona, 5, 14, 768)
    check(BaselineMona, 3, 16, 1024)
    timing()
