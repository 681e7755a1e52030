"""Mona adapter on the B200 kernels — host-side mirror of the reference's src/adapters/mona.py.

Kept identical to the reference (so it is a drop-in and checkpoints interchange):
  * class names, constructor signatures, forward(x, hw_shapes) call convention ([N,B,D] in/out for the
    adapter, [B,N,D] for BatchFirstMonaWrapper) — reference mona.py:38-67, :96-151
  * parameter names / shapes / creation order / initialisers: project1, project2, adapter_conv.conv{1,2,3},
    adapter_conv.projector, norm, gamma (1e-6), gammax (1) — mona.py:78-83, :104-113; the modules are
    real nn.Linear / nn.Conv2d / nn.LayerNorm instances used purely as parameter containers, so the CPU
    RNG stream consumed at construction and the state-dict keys match the reference bit for bit.
  * inject_mona_variant_to_{clip,open_clip}(model, variant, bottleneck_dim, num_layers) -> (model, count)
    — mona.py:495-575, :578-680 (attaches `block.mona`, wraps the *instance* forward, passes **kwargs).
Different: forward never touches ATen math.  It runs
  LN+mix kernel -> tcgen05 GEMM (project1) -> one-CTA-per-image stencil/projector/GELU/dropout kernel
  -> tcgen05 GEMM (project2, residual fused)   and a hand-written backward (see MonaFunction).
"""
import math

import torch
import torch.nn as nn

from .. import _lib as L
from .. import ops

_LN_EPS = 1e-5  # nn.LayerNorm default used by the reference adapter (mona.py:111)


def _next_seed():
    # dropout seeds come from torch's CPU generator so torch.manual_seed() makes runs reproducible
    return int(torch.randint(0, 2 ** 62, (1,)).item())


def _shadow(owner, w, dt, transpose):
    """Low-precision (optionally transposed) copy of a projection weight: taken from the trainer's one-launch
    `ops.CastPlan` when it is fresh for this parameter, converted on the spot otherwise."""
    plan = getattr(owner, "_ngu_castplan", None) if owner is not None else None
    if plan is not None and plan.fresh and plan.dtype == dt:
        t = plan.get(owner.project1.weight if w is owner.project1.weight or w.data_ptr() == owner.project1.weight.data_ptr()
                     else owner.project2.weight, transpose)
        if t is not None:
            return t
    return ops.cast(w, dt, transpose=transpose)


class MonaFunction(torch.autograd.Function):
    """y = x + project2(dropout(gelu(convstage(project1(LN(x)*gamma + x*gammax)))))  on [B,N,D].
    Variant tensors (freq_filter, noise-estimator weights) are None for the baseline adapter."""

    @staticmethod
    def forward(ctx, x, norm_w, norm_b, gamma, gammax, w1, b1, k3, b3, k5, b5, k7, b7, pw, pb, w2, b2,
                freq, ne_w1, ne_b1, ne_w2, ne_b2, hw, has_cls, drop_p, seed, owner=None):
        B, N, D = x.shape
        C = w1.shape[0]
        x2 = x.contiguous().view(B * N, D)
        dt = x2.dtype
        u, mean, rstd = ops.ln_fwd(x2, norm_w, norm_b, _LN_EPS, gamma=gamma, gammax=gammax)
        h = ops.gemm(u, _shadow(owner, w1, dt, False), bias=b1)                      # [M, C]
        conv_w = (k3, b3, k5, b5, k7, b7, pw, pb, freq, ne_w1, ne_b1, ne_w2, ne_b2)
        g = ops.mona_conv_fwd(h.view(B, N, C), conv_w, hw, has_cls, drop_p, seed)     # [B, N, C]
        y = ops.gemm(g.view(B * N, C), _shadow(owner, w2, dt, False), bias=b2, aux=x2, aux_mode=L.AUX_RESIDUAL)
        ctx.save_for_backward(x2, mean, rstd, u, h, g, norm_w, norm_b, gamma, gammax, w1, k3, b3, k5, b5, k7, b7, pw, pb, w2,
                              *[t for t in (freq, ne_w1, ne_b1, ne_w2, ne_b2) if t is not None])
        ctx.variant = (freq is not None, ne_w1 is not None)
        ctx.owner = owner
        ctx.meta = (B, N, D, C, hw, has_cls, drop_p, seed)
        return y.view(B, N, D)

    @staticmethod
    def backward(ctx, dy):
        sv = ctx.saved_tensors
        (x2, mean, rstd, u, h, g, norm_w, norm_b, gamma, gammax, w1, k3, b3, k5, b5, k7, b7, pw, pb, w2) = sv[:20]
        has_freq, has_noise = ctx.variant
        extra = list(sv[20:])
        freq = extra.pop(0) if has_freq else None
        ne_w1, ne_b1, ne_w2, ne_b2 = (extra if has_noise else [None] * 4)
        B, N, D, C, hw, has_cls, drop_p, seed = ctx.meta
        dev, dt = x2.device, x2.dtype
        dy2 = dy.contiguous().view(B * N, D)
        # Gradient sink: when the trainer has pre-allocated flat fp32 .grad buffers for this adapter's parameters
        # (dp.GradBuckets), the kernels accumulate straight into them (they all accumulate with += / atomics) and
        # autograd gets None -> no per-parameter zero-fill and add kernels.
        owner = ctx.owner
        sink = None
        if owner is not None:
            ps = list(owner.parameters())
            if ps and all(getattr(p, "_ngu_sink", None) is not None and p.grad is not None for p in ps):
                sink = ps[0]._ngu_sink
        c = owner.adapter_conv if owner is not None else None

        def buf(param_attr, shape):
            if sink is not None:
                return param_attr.grad
            return torch.zeros(*shape, device=dev, dtype=torch.float32)

        # project2:  y = x + g W2^T + b2
        dg = ops.gemm(dy2, _shadow(owner, w2, dt, True))                              # [M, C] = dy W2
        g2 = g.view(B * N, C)
        dw2 = buf(owner.project2.weight if sink else None, (D, C))
        ops.wgrad(dy2, g2, out=dw2)                                                   # [D, C]
        # conv stage (recomputes z / a from h), also yields d project1.bias
        if sink is not None:
            dk3, db3, dk5, db5, dk7, db7 = (c.conv1.weight.grad, c.conv1.bias.grad, c.conv2.weight.grad, c.conv2.bias.grad,
                                            c.conv3.weight.grad, c.conv3.bias.grad)
            dpw, dpb, db1 = c.projector.weight.grad, c.projector.bias.grad, owner.project1.bias.grad
            vt = c.variant_tensors()
            dfreq = vt[0].grad if has_freq else None
            dn = [t.grad for t in vt[1:]] if has_noise else [None] * 4
        else:
            z = lambda *s: torch.zeros(*s, device=dev, dtype=torch.float32)
            dk3, db3, dk5, db5, dk7, db7 = z(*k3.shape), z(C), z(*k5.shape), z(C), z(*k7.shape), z(C)
            dpw, dpb, db1 = z(*pw.shape), z(C), z(C)
            dfreq = z(C) if has_freq else None
            dn = [z(*t.shape) for t in (ne_w1, ne_b1, ne_w2, ne_b2)] if has_noise else [None] * 4
        dh = ops.mona_conv_bwd(h.view(B, N, C), dg.view(B, N, C), (k3, b3, k5, b5, k7, b7, pw, pb, freq, ne_w1, ne_b1, ne_w2, ne_b2),
                               (dk3, db3, dk5, db5, dk7, db7, dpw, dpb, db1, dfreq, *dn), hw, has_cls, drop_p, seed)
        dh2 = dh.view(B * N, C)
        # project1:  h = u W1^T + b1
        dw1t = ops.wgrad(u, dh2)                                                      # [D, C] = dW1^T
        du = ops.gemm(dh2, _shadow(owner, w1, dt, True))                              # [M, D] = dh W1
        # LN mix + residual
        if sink is not None:
            dnw, dnb, dgam, dgamx, db2 = owner.norm.weight.grad, owner.norm.bias.grad, owner.gamma.grad, owner.gammax.grad, owner.project2.bias.grad
        else:
            dnw, dnb, dgam, dgamx, db2 = (torch.zeros(D, device=dev, dtype=torch.float32) for _ in range(5))
        dx = ops.mona_pre_bwd(du, dy2, x2, mean, rstd, norm_w, norm_b, gamma, gammax, dnw, dnb, dgam, dgamx, db2)
        if sink is not None:
            owner.project1.weight.grad.add_(dw1t.t())
            buckets, key, count = sink
            buckets.notify(key, count)
            return (dx.view(B, N, D),) + (None,) * 26
        return (dx.view(B, N, D), dnw, dnb, dgam, dgamx, dw1t.t().contiguous(), db1, dk3, db3, dk5, db5, dk7, db7, dpw, dpb, dw2, db2,
                dfreq, dn[0], dn[1], dn[2], dn[3], None, None, None, None, None)


def _fused_ok(x, owner, C, hw, N, freq, ne_w1):
    """The fused bf16 path (csrc/mona_fused.cu): baseline / freq_enhanced stage, bottleneck 64, grids up to 16x16."""
    import os
    return (owner is not None and x.dtype == torch.bfloat16 and C == 64 and ne_w1 is None and hw[0] <= 16 and hw[1] <= 16
            and N <= 256 and x.shape[-1] % 128 == 0 and os.environ.get("NGU_MONA_FUSED", "1") != "0")


class MonaFusedFunction(torch.autograd.Function):
    """Same function as MonaFunction on the fused kernels: the LayerNorm mix is folded into project1 (x is consumed raw by
    TMA, u never exists), the stage runs out of shared memory in the same kernel, and backward is
    dy -> dg (GEMM) -> stage backward -> dx = dy + [dh | dh rstd][Wb; Wa] + beta_r x + alpha_r (GEMM epilogue), with every
    parameter gradient of the input mix derived from one token reduction G = x^T [dh | dh rstd]."""

    @staticmethod
    def forward(ctx, x, norm_w, norm_b, gamma, gammax, w1, b1, k3, b3, k5, b5, k7, b7, pw, pb, w2, b2, freq,
                hw, has_cls, drop_p, seed, owner):
        B, N, D = x.shape
        x = x.contiguous()
        x2 = x.view(B * N, D)
        der = ops.mona_derived(owner)
        h, hA, g, mean, rstd = ops.mona_fwd_stage(x, der, hw, has_cls, drop_p, seed, _LN_EPS)
        y = ops.gemm(g.view(B * N, 64), der.view("w2", torch.bfloat16, (D, 64)), bias=b2.detach(), aux=x2, aux_mode=L.AUX_RESIDUAL)
        ctx.save_for_backward(x2, h, hA, g, mean, rstd)
        ctx.owner, ctx.der = owner, der
        ctx.meta = (B, N, D, hw, has_cls, drop_p, seed, freq is not None)
        return y.view(B, N, D)

    @staticmethod
    def backward(ctx, dy):
        x2, h, hA, g, mean, rstd = ctx.saved_tensors
        owner, der = ctx.owner, ctx.der
        B, N, D, hw, has_cls, drop_p, seed, has_freq = ctx.meta
        dev = x2.device
        c = owner.adapter_conv
        dy2 = dy.contiguous().view(B * N, D)
        ps = list(owner.parameters())
        sink = ps[0]._ngu_sink if (ps and all(getattr(p, "_ngu_sink", None) is not None and p.grad is not None for p in ps)) else None
        names = ["norm.weight", "norm.bias", "gamma", "gammax", "project1.weight", "project1.bias", "conv1.weight", "conv1.bias",
                 "conv2.weight", "conv2.bias", "conv3.weight", "conv3.bias", "projector.weight", "projector.bias",
                 "project2.weight", "project2.bias"] + (["freq"] if has_freq else [])
        params = [owner.norm.weight, owner.norm.bias, owner.gamma, owner.gammax, owner.project1.weight, owner.project1.bias,
                  c.conv1.weight, c.conv1.bias, c.conv2.weight, c.conv2.bias, c.conv3.weight, c.conv3.bias,
                  c.projector.weight, c.projector.bias, owner.project2.weight, owner.project2.bias] + ([c.freq_filter] if has_freq else [])
        gr = {n: (p.grad if sink is not None else torch.zeros(p.shape, device=dev, dtype=torch.float32)) for n, p in zip(names, params)}
        # project2: y = x + g W2^T + b2
        dg = ops.gemm(dy2, der.view("w2_t", torch.bfloat16, (64, D)))                  # [M, 64] = dy W2
        ops.wgrad(dy2, g.view(B * N, 64), out=gr["project2.weight"])                   # [D, 64]
        ops.colsum(dy2, out=gr["project2.bias"])
        # stage backward -> [dh | dh rstd], row terms of the LayerNorm backward, stage reductions into ws
        dhcat, rowab, ws = ops.mona_bwd_stage(h, hA, dg.view(B, N, 64), mean, rstd, der, D, hw, has_cls, drop_p, seed,
                                              gr["projector.weight"], gr["projector.bias"])
        ops.wgrad(x2, dhcat, out=ws[:D * 128].view(D, 128))                            # G = x^T [dh | dh rstd]
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.gemm(dhcat, der.view("wcat_t", torch.bfloat16, (D, 128)), aux=dy2, aux_mode=L.AUX_MONA_DX, aux2=x2, rowab=rowab)
            dx = dx.view(B, N, D)
        ops.mona_finish(owner, {"dw1": gr["project1.weight"], "db1": gr["project1.bias"], "dln_w": gr["norm.weight"],
                                "dln_b": gr["norm.bias"], "dgamma": gr["gamma"], "dgammax": gr["gammax"],
                                "dk3": gr["conv1.weight"], "db3": gr["conv1.bias"], "dk5": gr["conv2.weight"], "db5": gr["conv2.bias"],
                                "dk7": gr["conv3.weight"], "db7": gr["conv3.bias"], "dfreq": gr.get("freq")}, ws, D)
        if sink is not None:
            buckets, key, count = sink
            buckets.notify(key, count)
            return (dx,) + (None,) * 22
        return (dx,) + tuple(gr[n] for n in names[:16]) + (gr.get("freq"),) + (None,) * 5


class _MonaOpBase(nn.Module):
    """Parameter container for the multi-scale depthwise stage.  forward() is not used by the adapters (the fused
    stage kernel consumes the parameters directly)."""

    def _build_convs(self, in_features):
        self.conv1 = nn.Conv2d(in_features, in_features, kernel_size=3, padding=1, groups=in_features)
        self.conv2 = nn.Conv2d(in_features, in_features, kernel_size=5, padding=2, groups=in_features)
        self.conv3 = nn.Conv2d(in_features, in_features, kernel_size=7, padding=3, groups=in_features)
        self.projector = nn.Conv2d(in_features, in_features, kernel_size=1)

    def _build_noise_estimator(self, in_features):
        self.noise_estimator = nn.Sequential(
            nn.AdaptiveAvgPool2d(1),
            nn.Conv2d(in_features, in_features // 4, 1),
            nn.ReLU(inplace=True),
            nn.Conv2d(in_features // 4, 3, 1),
            nn.Softmax(dim=1),
        )

    def variant_tensors(self):
        """(freq_filter, ne_w1, ne_b1, ne_w2, ne_b2) with None for absent parts."""
        freq = getattr(self, "freq_filter", None)
        ne = getattr(self, "noise_estimator", None)
        if ne is None:
            return freq, None, None, None, None
        return freq, ne[1].weight, ne[1].bias, ne[3].weight, ne[3].bias


class BaselineMonaOp(_MonaOpBase):
    """reference mona.py:75-93"""

    def __init__(self, in_features):
        super().__init__()
        self._build_convs(in_features)


class NoiseAwareMonaOp(_MonaOpBase):
    """reference mona.py:159-196: per-image softmax weights of the three depthwise branches."""

    def __init__(self, in_features):
        super().__init__()
        self._build_convs(in_features)
        self._build_noise_estimator(in_features)


class FreqEnhancedMonaOp(_MonaOpBase):
    """reference mona.py:261-296: learnable per-channel frequency filter (rfft2 -> x f_c -> irfft2 == scale by f_c)."""

    def __init__(self, in_features):
        super().__init__()
        self._build_convs(in_features)
        self.freq_filter = nn.Parameter(torch.ones(in_features))


class HybridNoiseFreqMonaOp(_MonaOpBase):
    """reference mona.py:370-425: frequency filter, then noise-aware branch weights."""

    def __init__(self, in_features):
        super().__init__()
        self._build_convs(in_features)
        self.freq_filter = nn.Parameter(torch.ones(in_features))
        self._build_noise_estimator(in_features)


class BaselineMona(nn.Module):
    """Baseline Mona adapter.  forward(x [N,B,D], hw_shapes) -> [N,B,D]   (reference mona.py:96-151).
    The variants below differ only in the `adapter_conv` stage."""

    op_class = BaselineMonaOp

    def __init__(self, in_dim, bottleneck_dim=64):
        super().__init__()
        self.project1 = nn.Linear(in_dim, bottleneck_dim)
        self.nonlinear = torch.nn.functional.gelu  # attribute kept for parity; the kernel applies exact-erf GELU
        self.project2 = nn.Linear(bottleneck_dim, in_dim)
        self.dropout = nn.Dropout(p=0.1)
        self.adapter_conv = self.op_class(bottleneck_dim)
        self.norm = nn.LayerNorm(in_dim)
        self.gamma = nn.Parameter(torch.ones(in_dim) * 1e-6)
        self.gammax = nn.Parameter(torch.ones(in_dim))

    def forward_batch_first(self, xb, hw_shapes=None):
        """[B,N,D] -> [B,N,D]; the layout the kernels work in."""
        B, N, D = xb.shape
        if hw_shapes is not None:
            hw, has_cls = (int(hw_shapes[0]), int(hw_shapes[1])), True
        else:  # all tokens form a sqrt(n) x sqrt(n) grid, no CLS (reference mona.py:140-144)
            s = int(math.sqrt(N))
            hw, has_cls = (s, s), False
        p = self.dropout.p if (self.training and self.dropout.p > 0) else 0.0
        seed = _next_seed() if p > 0 else 0
        c = self.adapter_conv
        freq, ne_w1, ne_b1, ne_w2, ne_b2 = c.variant_tensors()
        if _fused_ok(xb, self, self.project1.weight.shape[0], hw, N, freq, ne_w1):
            return MonaFusedFunction.apply(xb, self.norm.weight, self.norm.bias, self.gamma, self.gammax,
                                           self.project1.weight, self.project1.bias,
                                           c.conv1.weight, c.conv1.bias, c.conv2.weight, c.conv2.bias, c.conv3.weight, c.conv3.bias,
                                           c.projector.weight, c.projector.bias, self.project2.weight, self.project2.bias, freq,
                                           hw, has_cls, p, seed, self)
        return MonaFunction.apply(xb, self.norm.weight, self.norm.bias, self.gamma, self.gammax,
                                  self.project1.weight, self.project1.bias,
                                  c.conv1.weight, c.conv1.bias, c.conv2.weight, c.conv2.bias, c.conv3.weight, c.conv3.bias,
                                  c.projector.weight, c.projector.bias,
                                  self.project2.weight, self.project2.bias,
                                  freq, ne_w1, ne_b1, ne_w2, ne_b2, hw, has_cls, p, seed, self)

    def forward(self, x, hw_shapes=None):
        return self.forward_batch_first(x.permute(1, 0, 2), hw_shapes).permute(1, 0, 2)


class NoiseAwareMona(BaselineMona):
    """reference mona.py:198-258"""
    op_class = NoiseAwareMonaOp


class FreqEnhancedMona(BaselineMona):
    """reference mona.py:298-367 (the default --mona_variant of src/models/biomedclip/finetune.py:76)"""
    op_class = FreqEnhancedMonaOp


class HybridNoiseFreqMona(BaselineMona):
    """reference mona.py:427-492 (what scripts/biomedclip.sh uses)"""
    op_class = HybridNoiseFreqMonaOp


class BatchFirstMonaWrapper(nn.Module):
    """[B,N,D] adapter for open_clip-style trunks; attribute name `clip_mona` is part of the checkpoint
    key format (reference mona.py:38-67)."""

    def __init__(self, mona_adapter):
        super().__init__()
        self.clip_mona = mona_adapter

    def forward(self, x, hw_shapes=None):
        m = self.clip_mona
        if hasattr(m, "forward_batch_first"):
            return m.forward_batch_first(x, hw_shapes)  # skip the two cancelling permutes
        return m(x.permute(1, 0, 2), hw_shapes).permute(1, 0, 2)


_VARIANTS = {"baseline": BaselineMona, "noise_aware": NoiseAwareMona, "freq_enhanced": FreqEnhancedMona, "hybrid": HybridNoiseFreqMona}


def register_variant(name, cls):
    _VARIANTS[name] = cls


def _variant_class(variant):
    if variant not in _VARIANTS:
        raise ValueError(f"Unknown variant: {variant}. Choose from {list(_VARIANTS.keys())}")
    return _VARIANTS[variant]


def _wrap_block_forward(block, hw_shapes):
    inner = block.forward

    def forward_with_mona(x, **kwargs):
        return block.mona(inner(x, **kwargs), hw_shapes)

    block.forward = forward_with_mona


def inject_mona_variant_to_clip(model, variant="hybrid", bottleneck_dim=64, num_layers=None):
    """OpenAI-CLIP layout (visual.transformer.resblocks, [N,B,D]); reference mona.py:495-575."""
    cls = _variant_class(variant)
    count = 0
    visual = getattr(model, "visual", None)
    if visual is not None and hasattr(visual, "transformer"):
        tr = visual.transformer
        if hasattr(visual, "input_resolution"):
            res = visual.input_resolution
        elif hasattr(visual, "image_size"):
            res = visual.image_size[0] if isinstance(visual.image_size, tuple) else visual.image_size
        else:
            raise AttributeError("Model does not have input_resolution or image_size attribute")
        grid = res // visual.conv1.kernel_size[0]
        if hasattr(tr, "resblocks"):
            blocks = tr.resblocks
            n = len(blocks) if num_layers is None else min(num_layers, len(blocks))
            for i in range(n):
                blocks[i].mona = cls(tr.width, bottleneck_dim)
                _wrap_block_forward(blocks[i], (grid, grid))
                count += 1
    print(f"✓ Injected {variant} MONA adapters to {count} layers (OpenAI CLIP vision encoder)")
    return model, count


def inject_mona_variant_to_open_clip(model, variant="hybrid", bottleneck_dim=64, num_layers=None):
    """open_clip layout (visual.trunk.blocks for timm towers, visual.transformer.resblocks otherwise), [B,N,D];
    reference mona.py:578-680."""
    cls = _variant_class(variant)
    count = 0
    visual = getattr(model, "visual", None)
    if visual is not None:
        blocks = dim = hw = None
        if hasattr(visual, "trunk"):
            trunk = visual.trunk
            dim = trunk.embed_dim
            g = int(math.sqrt(trunk.patch_embed.num_patches))
            hw = (g, g)
            blocks = getattr(trunk, "blocks", None)
        elif hasattr(visual, "transformer"):
            dim = visual.transformer.width
            if hasattr(visual, "grid_size"):
                hw = (visual.grid_size[0], visual.grid_size[0])
            elif hasattr(visual, "image_size") and hasattr(visual, "patch_size"):
                im = visual.image_size[0] if isinstance(visual.image_size, tuple) else visual.image_size
                ps = visual.patch_size[0] if isinstance(visual.patch_size, tuple) else visual.patch_size
                hw = (im // ps, im // ps)
            blocks = getattr(visual.transformer, "resblocks", None)
        if blocks is not None and dim is not None:
            n = len(blocks) if num_layers is None else min(num_layers, len(blocks))
            for i in range(n):
                blocks[i].mona = BatchFirstMonaWrapper(cls(dim, bottleneck_dim))
                _wrap_block_forward(blocks[i], hw)
                count += 1
    print(f"✓ Injected {variant} MONA adapters to {count} layers (open_clip vision encoder)")
    return model, count
