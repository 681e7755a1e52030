"""BiomedCLIP-shaped model (open_clip CustomTextCLIP layout) assembled on the B200 kernels.

open_clip / timm / transformers' BERT are un-vendored pinned dependencies of the reference
(`create_model_from_pretrained("hf-hub:microsoft/BiomedCLIP-PubMedBERT_256-vit_base_patch16_224")`,
src/models/biomedclip/finetune.py:116).  This module restates the *structure* the reference touches:
    model.visual.trunk (timm ViT-B/16), model.visual.head.proj (768->512, no bias),
    model.text.transformer (BERT-base: embeddings / encoder.layer[i].attention.{self.{query,key,value},
    output.{dense,LayerNorm}} / intermediate.dense / output.{dense,LayerNorm}), model.text.proj
    (CLS pooler + MLP 768->640->GELU->512, no biases), encode_image / encode_text, logit_scale.
Weights are synthetic unless loaded from a state dict with these (open_clip) key names.
"""
import math

import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .linear import frozen_copies
from .linear import Proj, proj_fwd
from .vit import VisionTransformer, HeadFunction


class _Head(nn.Module):
    def __init__(self, dim, out_dim):
        super().__init__()
        self.proj = nn.Linear(dim, out_dim, bias=False)


class TimmVisual(nn.Module):
    """open_clip TimmModel: trunk + head (pool = CLS token, proj = linear without bias)."""

    def __init__(self, embed_dim=512, **vit_kw):
        super().__init__()
        self.trunk = VisionTransformer(**vit_kw)
        self.head = _Head(self.trunk.embed_dim, embed_dim)

    def forward(self, images):
        x = self.trunk.forward_features(images)
        return HeadFunction.apply(x, self.trunk.norm, self.head.proj)


# ---- BERT-shaped parameter containers (HF attribute names; forward is the fused kernel pipeline below) ----
class _BertSelfAttention(nn.Module):
    def __init__(self, d, heads):
        super().__init__()
        self.num_attention_heads = heads
        self.query, self.key, self.value = nn.Linear(d, d), nn.Linear(d, d), nn.Linear(d, d)


class _BertSelfOutput(nn.Module):
    def __init__(self, d, eps):
        super().__init__()
        self.dense = nn.Linear(d, d)
        self.LayerNorm = nn.LayerNorm(d, eps=eps)


class _BertAttention(nn.Module):
    def __init__(self, d, heads, eps):
        super().__init__()
        self.self = _BertSelfAttention(d, heads)
        self.output = _BertSelfOutput(d, eps)


class _BertIntermediate(nn.Module):
    def __init__(self, d, dm):
        super().__init__()
        self.dense = nn.Linear(d, dm)


class _BertOutput(nn.Module):
    def __init__(self, d, dm, eps):
        super().__init__()
        self.dense = nn.Linear(dm, d)
        self.LayerNorm = nn.LayerNorm(d, eps=eps)


class _BertLayer(nn.Module):
    def __init__(self, d, heads, dm, eps):
        super().__init__()
        self.attention = _BertAttention(d, heads, eps)
        self.intermediate = _BertIntermediate(d, dm)
        self.output = _BertOutput(d, dm, eps)


class _BertEncoder(nn.Module):
    def __init__(self, layers, d, heads, dm, eps):
        super().__init__()
        self.layer = nn.ModuleList([_BertLayer(d, heads, dm, eps) for _ in range(layers)])


class _BertEmbeddings(nn.Module):
    def __init__(self, vocab, d, max_pos, eps):
        super().__init__()
        self.word_embeddings = nn.Embedding(vocab, d, padding_idx=0)
        self.position_embeddings = nn.Embedding(max_pos, d)
        self.token_type_embeddings = nn.Embedding(2, d)
        self.LayerNorm = nn.LayerNorm(d, eps=eps)


class BertModel(nn.Module):
    def __init__(self, vocab=30522, d=768, layers=12, heads=12, dm=3072, max_pos=512, eps=1e-12):
        super().__init__()
        self.embeddings = _BertEmbeddings(vocab, d, max_pos, eps)
        self.encoder = _BertEncoder(layers, d, heads, dm, eps)


class _PackedAttnFunction(torch.autograd.Function):
    """softmax(q k^T / sqrt(dh)) v on a packed [B*S, 3*D] projection with the key-padding lengths of the batch."""

    @staticmethod
    def forward(ctx, qkv, B, S, H, kv_len):
        qkv = qkv.contiguous()
        dh = qkv.shape[1] // (3 * H)
        o, lse = ops.attn_fwd_packed(qkv, B, S, H, dh, kv_len=kv_len)
        ctx.save_for_backward(qkv, o, lse)
        ctx.meta = (B, S, H, dh, kv_len)
        return o

    @staticmethod
    def backward(ctx, do):
        qkv, o, lse = ctx.saved_tensors
        B, S, H, dh, kv_len = ctx.meta
        return ops.attn_bwd_packed(qkv, o, lse, do.contiguous(), B, S, H, dh, kv_len=kv_len), None, None, None, None


class _TextProjFunction(torch.autograd.Function):
    """proj[2](gelu(proj[0](cls))) with frozen, bias-free weights (open_clip `mlp` text projection)."""

    @staticmethod
    def forward(ctx, cls, fc, out):
        c2 = cls.contiguous()
        p1, p2 = Proj(fc, c2.dtype), Proj(out, c2.dtype)
        (a, der), _ = proj_fwd(c2, p1, act=L.ACT_GELU, save_pre=True)
        y, _ = proj_fwd(a, p2)
        ctx.save_for_backward(der)
        ctx.p = (p1, p2)
        return y

    @staticmethod
    def backward(ctx, dy):
        (der,) = ctx.saved_tensors
        p1, p2 = ctx.p
        dpre = ops.gemm(dy.contiguous(), p2.WT, aux=der, aux_mode=L.AUX_DACT)
        return ops.gemm(dpre, p1.WT), None, None


class HFTextEncoder(nn.Module):
    """open_clip HFTextEncoder with `cls_last_hidden_state_pooler` and `mlp` projection (BiomedCLIP config)."""

    def __init__(self, embed_dim=512, d=768, pad_token_id=0, **bert_kw):
        super().__init__()
        self.transformer = BertModel(d=d, **bert_kw)
        hidden = (d + embed_dim) // 2
        self.proj = nn.Sequential(nn.Linear(d, hidden, bias=False), nn.GELU(), nn.Linear(hidden, embed_dim, bias=False))
        self.pad_token_id = pad_token_id
        self.compute_dtype = torch.bfloat16
        self._pad_flag = None

    def check_padding(self):
        """Raise if any batch seen since the last check was not right-padded (host read of the device flag)."""
        if self._pad_flag is not None and int(self._pad_flag.item()) != 0:
            self._pad_flag.zero_()
            raise NotImplementedError("ngu B200 path: only right-padded token batches (a suffix of pad ids) are supported")

    def _qkv_weights(self, sa, dt):
        """Fused [3d, d] q/k/v weight + bias (frozen), cached on the module."""
        key = (dt, sa.query.weight.device, sa.query.weight._version, sa.key.weight._version, sa.value.weight._version)
        c = getattr(sa, "_ngu_qkv", None)
        if c is None or c[0] != key:
            w = torch.cat([sa.query.weight, sa.key.weight, sa.value.weight], 0).detach().float().contiguous()
            b = torch.cat([sa.query.bias, sa.key.bias, sa.value.bias], 0).detach().float().contiguous()
            c = (key, ops.cast(w, dt), b)
            sa._ngu_qkv = c
        return c[1], c[2]

    def forward(self, ids):
        """ids int64 [B,S] -> [B, embed_dim].  Frozen tower (src/models/biomedclip/finetune.py:166-167, the default): one
        fused forward-only pipeline without an autograd graph.  With LoRA on the BERT projections
        (inject_lora_to_biomedclip(tune_text_encoder=True), reference lora.py:317-367) the layers run as autograd nodes."""
        # Key-padding mask (open_clip HFTextEncoder.forward: attn_mask = (x != pad_token_id)): tokenizers pad on the right,
        # so the mask is a per-sequence valid length; padded keys are masked in every layer, padded query rows compute
        # values nobody reads (CLS pooling).  Any other mask shape is refused loudly.
        # The lengths are computed on the device every call (one tiny kernel, no host round trip); the right-padding
        # property lands in a device flag that is checked eagerly in eval mode (a host sync, inference / tests) and lazily
        # through check_padding() in train mode, where the step must not synchronise (finetune.py:272-303 hot loop).
        if ids.is_cuda:
            if self._pad_flag is None or self._pad_flag.device != ids.device:
                self._pad_flag = torch.zeros(1, device=ids.device, dtype=torch.int32)
            kv_len = ops.kv_len(ids.contiguous(), self.pad_token_id, self._pad_flag)
            if not self.training:
                self.check_padding()
        else:
            raise L.NguError("ngu ops need CUDA tensors: there is no CPU fallback in the product path")
        has_lora = any(getattr(m, "r", 0) and hasattr(m, "w_lora_A") for m in self.modules())
        if has_lora or any(p.requires_grad for p in self.parameters()):
            return self._forward_trainable(ids, kv_len)
        with torch.no_grad():
            return self._forward_frozen(ids, kv_len)

    def _embed(self, ids, dt):
        emb = self.transformer.embeddings
        for p in emb.parameters():
            if p.requires_grad:
                raise NotImplementedError("ngu B200 path: BERT embeddings must stay frozen")
        with torch.no_grad():
            x = ops.embed_tokens(ids.contiguous(), emb.word_embeddings.weight.detach(), emb.position_embeddings.weight.detach(),
                                 emb.token_type_embeddings.weight.detach()[0].contiguous(), dt)
            x, _, _ = ops.ln_fwd(x, emb.LayerNorm.weight.detach(), emb.LayerNorm.bias.detach(), emb.LayerNorm.eps, save_stats=False)
        return x

    def _forward_trainable(self, ids, kv_len):
        """Post-LN BERT layers as autograd nodes over the same kernels: LinearLoRA / frozen projections, packed attention
        with the key-padding lengths, LayerNorm, GELU MLP with the residual in the GEMM epilogue."""
        from .adapters.lora import _apply_linear
        from .openai_clip import _LnFunction, _MlpResidualFunction
        for n, p in self.named_parameters():
            if p.requires_grad and "lora" not in n:
                raise NotImplementedError(f"ngu B200 path: only LoRA factors of the text tower may be trainable (got {n})")
        tr = self.transformer
        dt = self.compute_dtype
        B, S = ids.shape
        x = self._embed(ids, dt).view(B, S, -1)
        d = x.shape[-1]
        for lyr in tr.encoder.layer:
            sa, so = lyr.attention.self, lyr.attention.output
            H = sa.num_attention_heads
            qkv = torch.cat([_apply_linear(sa.query, x), _apply_linear(sa.key, x), _apply_linear(sa.value, x)], -1)
            ao = _PackedAttnFunction.apply(qkv.view(B * S, 3 * d), B, S, H, kv_len)
            s1 = x + _apply_linear(so.dense, ao.view(B, S, d))
            x1 = _LnFunction.apply(s1, so.LayerNorm)
            s2 = _MlpResidualFunction.apply(x1, x1, lyr.intermediate.dense, lyr.output.dense, L.ACT_GELU)
            x = _LnFunction.apply(s2, lyr.output.LayerNorm)
        return _TextProjFunction.apply(x[:, 0, :], self.proj[0], self.proj[2])

    def _forward_frozen(self, ids, kv_len):
        tr = self.transformer
        dt = self.compute_dtype
        B, S = ids.shape
        x = self._embed(ids, dt)
        d = x.shape[-1]
        for lyr in tr.encoder.layer:
            sa = lyr.attention.self
            H = sa.num_attention_heads
            Wqkv, bqkv = self._qkv_weights(sa, dt)
            qkv = ops.gemm(x, Wqkv, bias=bqkv)
            ao, _ = ops.attn_fwd_packed(qkv, B, S, H, d // H, kv_len=kv_len)
            so = lyr.attention.output
            y = ops.gemm(ao, frozen_copies(so.dense.weight, dt)[0], bias=so.dense.bias.detach(), aux=x, aux_mode=L.AUX_RESIDUAL)
            x, _, _ = ops.ln_fwd(y, so.LayerNorm.weight.detach(), so.LayerNorm.bias.detach(), so.LayerNorm.eps, save_stats=False)
            h = ops.gemm(x, frozen_copies(lyr.intermediate.dense.weight, dt)[0], bias=lyr.intermediate.dense.bias.detach(), act=L.ACT_GELU)
            y = ops.gemm(h, frozen_copies(lyr.output.dense.weight, dt)[0], bias=lyr.output.dense.bias.detach(), aux=x, aux_mode=L.AUX_RESIDUAL)
            x, _, _ = ops.ln_fwd(y, lyr.output.LayerNorm.weight.detach(), lyr.output.LayerNorm.bias.detach(), lyr.output.LayerNorm.eps, save_stats=False)
        cls = x.view(B, S, d)[:, 0, :]                       # CLS pooling: strided rows feed the GEMM directly
        h = ops.gemm(cls, frozen_copies(self.proj[0].weight, dt)[0], act=L.ACT_GELU)
        return ops.gemm(h, frozen_copies(self.proj[2].weight, dt)[0])


class BiomedCLIP(nn.Module):
    """CustomTextCLIP-shaped container: .visual (TimmVisual), .text (HFTextEncoder), .logit_scale."""

    def __init__(self, embed_dim=512, vision=None, text=None):
        super().__init__()
        self.visual = TimmVisual(embed_dim=embed_dim, **(vision or {}))
        self.text = HFTextEncoder(embed_dim=embed_dim, **(text or {}))
        self.logit_scale = nn.Parameter(torch.ones([]) * math.log(1 / 0.07))
        self.context_length = 256

    def set_compute_dtype(self, dtype):
        """torch.bfloat16 = tcgen05 product path; torch.float32 = fp32 check mode (CUDA cores)."""
        self.visual.trunk.compute_dtype = dtype
        self.text.compute_dtype = dtype
        return self

    def encode_image(self, image, normalize: bool = False):
        f = self.visual(image)
        return nn.functional.normalize(f, dim=-1) if normalize else f

    def encode_text(self, text, normalize: bool = False):
        f = self.text(text)
        return nn.functional.normalize(f, dim=-1) if normalize else f


def init_synthetic_(model, seed=1, std=0.02):
    """Deterministic synthetic base weights: normal(0, std) for matrices/embeddings,
    LayerNorm = (1, 0), biases normal(0, std).  Call BEFORE injecting adapters."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("logit_scale"):
                continue
            if "norm" in name.lower() and name.endswith("weight") and p.dim() == 1:
                p.fill_(1.0)
            elif "norm" in name.lower() and name.endswith("bias"):
                p.zero_()
            else:
                p.copy_(torch.randn(p.shape, generator=g) * std)
    return model
