"""LoRA on the B200 kernels — host-side mirror of the reference's src/adapters/lora.py.

Same classes, constructor signatures, parameter names (`w_lora_A` [r,in], `w_lora_B` [out,r]),
initialisers and RNG consumption as the reference (lora.py:13-90): LinearLoRA builds a fresh
nn.Linear (burning the generator exactly like the reference does), copies the existing weights,
registers zero A/B, then kaiming-uniform(a=sqrt 5) on A; scaling = alpha / sqrt(r).  As in the
reference only `.weight` is frozen, so the bias of a wrapped projection stays trainable.
forward() is different: low-rank form, fused into the base projection's tcgen05 GEMM.
"""
import math

import torch
import torch.nn as nn

from .. import ops
from ..linear import Proj, proj_fwd, proj_bwd, require_causal_mask


class LoRALayer:
    def __init__(self, r: int, lora_alpha: int, dropout_rate: float = 0):
        self.r = r
        self.lora_alpha = lora_alpha
        self.dropout_rate = dropout_rate
        if self.r > 0:
            self.scaling = self.lora_alpha / math.sqrt(self.r)
        self.merged = False
        self.params_with_lora = {}

    def register_lora_param(self):
        for param_name, lora_name in self.params_with_lora.items():
            base = getattr(self, param_name)
            assert base.dim() == 2
            self.register_parameter(f"{lora_name}_lora_A", nn.Parameter(base.new_zeros((self.r, base.shape[1]))))
            self.register_parameter(f"{lora_name}_lora_B", nn.Parameter(base.new_zeros((base.shape[0], self.r))))
            base.requires_grad = False

    def init_lora_param(self):
        for _, lora_name in self.params_with_lora.items():
            if hasattr(self, f"{lora_name}_lora_A"):
                nn.init.kaiming_uniform_(getattr(self, f"{lora_name}_lora_A"), a=math.sqrt(5))
                nn.init.zeros_(getattr(self, f"{lora_name}_lora_B"))

    def merge_BA(self, param_name: str):
        lora_name = self.params_with_lora[param_name]
        return (getattr(self, f"{lora_name}_lora_B") @ getattr(self, f"{lora_name}_lora_A")).view(getattr(self, param_name).shape)


class _LoraLinearFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bias, A, B, module, seed):
        shp = x.shape
        x2 = x.contiguous().view(-1, shp[-1])
        pj = Proj(module, x2.dtype)
        y, saved = proj_fwd(x2, pj, seed=seed)
        ctx.pj, ctx.saved_lora, ctx.shape = pj, saved, shp
        return y.view(*shp[:-1], y.shape[-1])

    @staticmethod
    def backward(ctx, dy):
        pj = ctx.pj
        dy2 = dy.contiguous().view(-1, dy.shape[-1])
        dx, dbias, dA, dB = proj_bwd(dy2, pj, ctx.saved_lora, need_dx=ctx.needs_input_grad[0],
                                     need_bias=ctx.needs_input_grad[1])
        return (dx.view(ctx.shape) if dx is not None else None), dbias, dA, dB, None, None


class LinearLoRA(nn.Linear, LoRALayer):
    def __init__(self, existing_linear: nn.Linear, r: int = 0, lora_alpha: int = 1, dropout_rate: float = 0.0):
        super().__init__(in_features=existing_linear.in_features, out_features=existing_linear.out_features)
        self.load_state_dict(existing_linear.state_dict())
        LoRALayer.__init__(self, r=r, lora_alpha=lora_alpha, dropout_rate=dropout_rate)
        self.params_with_lora = {"weight": "w"}
        if r > 0:
            self.register_lora_param()
        self.init_lora_param()
        self.dropout = nn.Dropout(dropout_rate) if dropout_rate > 0 else None

    def forward(self, x: torch.Tensor):
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if (self.training and self.dropout is not None and self.dropout.p > 0) else 0
        A = getattr(self, "w_lora_A", None)
        B = getattr(self, "w_lora_B", None)
        return _LoraLinearFunction.apply(x, self.bias, A, B, self, seed)


class _SdpaFunction(torch.autograd.Function):
    """softmax(q k^T / sqrt(dh)) v on [L,B,D]-shaped projections (sequence-first), per head."""

    @staticmethod
    def forward(ctx, q, k, v, num_heads, causal):
        Lq, Bn, D = q.shape
        S = k.shape[0]
        dh = D // num_heads
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        st = ((D, Bn * D), (D, Bn * D), (D, Bn * D), (D, Bn * D))  # (batch stride, token stride) for [L,B,D]
        o, lse = ops.attn_fwd_strided(q, k, v, Bn, num_heads, Lq, S, dh, st, causal=causal)
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.meta = (Bn, num_heads, Lq, S, dh, st, causal)
        return o.view(Lq, Bn, D)

    @staticmethod
    def backward(ctx, do):
        q, k, v, o, lse = ctx.saved_tensors
        Bn, H, Lq, S, dh, st, causal = ctx.meta
        dq, dk, dv = ops.attn_bwd_strided(q, k, v, o, lse, do.contiguous(), Bn, H, Lq, S, dh, st, causal=causal)
        return dq, dk, dv, None, None


class PlainMultiheadAttentionLoRA(nn.Module):
    """nn.MultiheadAttention re-expressed as q/k/v/proj Linears with optional LoRA on each
    (reference lora.py:93-199).  Returns (out, None)."""

    def __init__(self, existing_mha: nn.MultiheadAttention, enable_lora: list = ["q", "k", "v", "o"], r: int = 0,
                 lora_alpha: int = 1, dropout_rate: float = 0.0):
        super().__init__()
        self.dropout = 0
        self.embed_dim = existing_mha.embed_dim
        self.kdim = existing_mha.kdim
        self.vdim = existing_mha.vdim
        self._qkv_same_embed_dim = existing_mha._qkv_same_embed_dim
        self.num_heads = existing_mha.num_heads
        self.batch_first = existing_mha.batch_first
        self.head_dim = existing_mha.head_dim
        E = self.embed_dim
        has_in_bias = existing_mha.in_proj_bias is not None
        self.q_proj = nn.Linear(E, E, bias=has_in_bias)
        self.k_proj = nn.Linear(E, E, bias=has_in_bias)
        self.v_proj = nn.Linear(E, E, bias=has_in_bias)
        self.proj = nn.Linear(E, E, bias=existing_mha.out_proj.bias is not None)
        with torch.no_grad():
            w = existing_mha.in_proj_weight.data
            for i, lin in enumerate((self.q_proj, self.k_proj, self.v_proj)):
                lin.weight.data.copy_(w[i * E:(i + 1) * E, :])
                if has_in_bias:
                    lin.bias.data.copy_(existing_mha.in_proj_bias.data[i * E:(i + 1) * E])
            self.proj.weight.data.copy_(existing_mha.out_proj.weight.data)
            if self.proj.bias is not None:
                self.proj.bias.data.copy_(existing_mha.out_proj.bias.data)
        for item in enable_lora:
            kw = dict(r=r, lora_alpha=lora_alpha, dropout_rate=dropout_rate)
            if item == "q":
                self.q_proj = LinearLoRA(self.q_proj, **kw)
            elif item == "k":
                self.k_proj = LinearLoRA(self.k_proj, **kw)
            elif item == "v":
                self.v_proj = LinearLoRA(self.v_proj, **kw)
            elif item == "o":
                self.proj = LinearLoRA(self.proj, **kw)

    def forward(self, query, key, value, key_padding_mask=None, need_weights=False, attn_mask=None, **kwargs):
        batched = query.dim() == 3
        if self.batch_first and batched:
            query, key, value = (t.transpose(1, 0) for t in (query, key, value))
        causal = False
        if attn_mask is not None:
            # the only mask on this path is CLIP's causal text mask (model.py build_attention_mask): verified, not assumed
            if query.shape[0] != key.shape[0]:
                raise NotImplementedError("attn_mask with Lq != S is not implemented on this path")
            require_causal_mask(attn_mask, query.shape[0])
            causal = True
        if key_padding_mask is not None:
            raise NotImplementedError("key_padding_mask is not used on the reference hot path")
        if (not causal and batched and query is key and key is value and self.head_dim == 64
                and query.dtype == torch.bfloat16 and query.transpose(0, 1).is_contiguous()):
            # Self-attention on a sequence-first VIEW of batch-first memory (what the CLIP vision tower hands its blocks,
            # model.py:250): project on the batch-first rows and pack q | k | v so the attention core runs on the tcgen05
            # kernels (attention_tc.cu / attention_long.cu) instead of the strided CUDA-core path; no layout copies.
            from ..biomedclip import _PackedAttnFunction
            xb = query.transpose(0, 1)                                  # [B, L, D] contiguous
            Bn, L, D = xb.shape
            qkv = torch.cat([_apply_linear(self.q_proj, xb), _apply_linear(self.k_proj, xb), _apply_linear(self.v_proj, xb)], -1)
            o = _PackedAttnFunction.apply(qkv.view(Bn * L, 3 * D), Bn, L, self.num_heads, None)
            o = _apply_linear(self.proj, o.view(Bn, L, D)).transpose(0, 1)
            if self.batch_first:
                return o.transpose(1, 0), None
            return o, None
        q = _apply_linear(self.q_proj, query)
        k = _apply_linear(self.k_proj, key)
        v = _apply_linear(self.v_proj, value)
        o = _SdpaFunction.apply(q, k, v, self.num_heads, causal)
        o = _apply_linear(self.proj, o)
        if self.batch_first and batched:
            return o.transpose(1, 0), None
        return o, None


class _FrozenLinearFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bias, module):
        shp = x.shape
        x2 = x.contiguous().view(-1, shp[-1])
        pj = Proj(module, x2.dtype)
        y, _ = proj_fwd(x2, pj)
        ctx.pj, ctx.shape = pj, shp
        return y.view(*shp[:-1], y.shape[-1])

    @staticmethod
    def backward(ctx, dy):
        dy2 = dy.contiguous().view(-1, dy.shape[-1])
        dx, dbias, _, _ = proj_bwd(dy2, ctx.pj, None, need_dx=ctx.needs_input_grad[0], need_bias=ctx.needs_input_grad[1])
        return (dx.view(ctx.shape) if dx is not None else None), dbias, None


def _apply_linear(lin, x):
    if isinstance(lin, LinearLoRA):
        return lin(x)
    return _FrozenLinearFunction.apply(x, lin.bias, lin)


def inject_lora_to_clip(model, lora_r=16, lora_alpha=32, lora_dropout=0.1, num_layers=None):
    """Swap nn.MultiheadAttention -> PlainMultiheadAttentionLoRA in visual.transformer.resblocks
    (reference lora.py:202-248)."""
    count = 0
    visual = getattr(model, "visual", None)
    if visual is not None and hasattr(visual, "transformer") and hasattr(visual.transformer, "resblocks"):
        blocks = visual.transformer.resblocks
        n = len(blocks) if num_layers is None else min(num_layers, len(blocks))
        for i in range(n):
            blk = blocks[i]
            if hasattr(blk, "attn") and isinstance(blk.attn, nn.MultiheadAttention):
                blk.attn = PlainMultiheadAttentionLoRA(blk.attn, enable_lora=["q", "k", "v", "o"], r=lora_r,
                                                       lora_alpha=lora_alpha, dropout_rate=lora_dropout)
                count += 1
    print(f"✓ Injected LoRA adapters to {count} layers (CLIP vision encoder)")
    return model, count


def _wrap(holder, name, kw):
    lin = getattr(holder, name, None)
    if isinstance(lin, nn.Linear):
        setattr(holder, name, LinearLoRA(lin, **kw))


def inject_lora_to_biomedclip(model, lora_r=16, lora_alpha=32, lora_dropout=0.1, num_layers=None, tune_text_encoder=False):
    """Wrap attn.qkv / attn.proj of visual.trunk.blocks (and optionally the BERT q/k/v/o) in LinearLoRA
    (reference lora.py:251-370)."""
    count = 0
    kw = dict(r=lora_r, lora_alpha=lora_alpha, dropout_rate=lora_dropout)
    visual = getattr(model, "visual", None)
    if visual is not None and hasattr(visual, "trunk") and hasattr(visual.trunk, "blocks"):
        blocks = visual.trunk.blocks
        n = len(blocks) if num_layers is None else min(num_layers, len(blocks))
        for i in range(n):
            attn = getattr(blocks[i], "attn", None)
            if attn is not None:
                _wrap(attn, "qkv", kw)
                _wrap(attn, "proj", kw)
                count += 1
    if tune_text_encoder:
        text = getattr(model, "text", None)
        enc = getattr(getattr(text, "transformer", None), "encoder", None)
        if enc is not None and hasattr(enc, "layer"):
            layers = enc.layer
            n = len(layers) if num_layers is None else min(num_layers, len(layers))
            for i in range(n):
                att = getattr(layers[i], "attention", None)
                if att is not None and hasattr(att, "self"):
                    for nm in ("query", "key", "value"):
                        _wrap(att.self, nm, kw)
                    if hasattr(att, "output"):
                        _wrap(att.output, "dense", kw)
                    count += 1
    print(f"✓ Injected LoRA adapters to {count} layers (BiomedCLIP vision encoder)")
    return model, count
