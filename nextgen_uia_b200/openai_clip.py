"""OpenAI-CLIP-shaped model on the B200 kernels — mirror of the reference's vendored
src/third_party/openai_clip/model.py (LayerNorm :163-169, QuickGELU :172-174, ResidualAttentionBlock :177-202,
Transformer :205-213, VisionTransformer :216-257, CLIP :260-391).

Same class names, constructor signatures and attribute / state-dict names (visual.conv1, visual.class_embedding,
visual.positional_embedding, visual.ln_pre, visual.transformer.resblocks.{i}.{attn.in_proj_weight, attn.in_proj_bias,
attn.out_proj, ln_1, mlp.c_fc, mlp.c_proj, ln_2}, visual.ln_post, visual.proj, transformer.*, token_embedding,
positional_embedding, ln_final, text_projection, logit_scale) so `inject_mona_variant_to_clip` / `inject_lora_to_clip`
and the reference checkpoints apply unchanged.  `block.attn` is a real nn.MultiheadAttention used as a parameter
container (the LoRA injection pattern-matches on that type, reference lora.py:237).

Blocks keep the reference call convention x [N,B,D] -> [N,B,D].  The vision tower hands the blocks a permuted VIEW of
batch-first memory exactly like the reference does (model.py:250), so the kernels (token-major [B,N,D]) run on it
with zero copies.
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L
from . import ops
from .linear import Proj, proj_fwd, proj_bwd, require_causal_mask
from .vit import BlockFunction, BlockSpec


class LayerNorm(nn.LayerNorm):
    """Parameter container; statistics are always fp32 in the kernel (the reference upcasts for fp16, model.py:166-169)."""


class QuickGELU(nn.Module):
    def forward(self, x):  # only reached outside the fused paths (API parity)
        raise NotImplementedError("QuickGELU runs inside the GEMM epilogue on the ngu B200 path")


class _PackedInProj:
    """nn.MultiheadAttention's packed in-projection seen as one Linear [3D, D] (q | k | v rows)."""

    def __init__(self, mha):
        self.weight, self.bias = mha.in_proj_weight, mha.in_proj_bias
        self.training = mha.training


class _LnFunction(torch.autograd.Function):
    """LayerNorm with frozen affine as a stand-alone autograd node (generic, non-fused block path)."""

    @staticmethod
    def forward(ctx, x, norm):
        shp = x.shape
        x2 = x.contiguous().view(-1, shp[-1])
        y, mean, rstd = ops.ln_fwd(x2, norm.weight.detach(), norm.bias.detach(), norm.eps)
        ctx.save_for_backward(x2, mean, rstd)
        ctx.norm, ctx.shape = norm, shp
        return y.view(shp)

    @staticmethod
    def backward(ctx, g):
        x2, mean, rstd = ctx.saved_tensors
        dx = ops.ln_bwd(g.contiguous().view(x2.shape), x2, mean, rstd, ctx.norm.weight.detach())
        return dx.view(ctx.shape), None


class _MlpResidualFunction(torch.autograd.Function):
    """y = res + c_proj(act(c_fc(h))) with frozen weights: two GEMMs, activation / derivative / residual in the epilogues."""

    @staticmethod
    def forward(ctx, h, res, fc, proj, act):
        shp = h.shape
        h2, r2 = h.contiguous().view(-1, shp[-1]), res.contiguous().view(-1, shp[-1])
        p1, p2 = Proj(fc, h2.dtype), Proj(proj, h2.dtype)
        (a, der), _ = proj_fwd(h2, p1, act=act, save_pre=True)
        y, _ = proj_fwd(a, p2, aux=r2, aux_mode=L.AUX_RESIDUAL)
        ctx.save_for_backward(der)
        ctx.p, ctx.shape = (p1, p2), shp
        return y.view(shp)

    @staticmethod
    def backward(ctx, dy):
        (der,) = ctx.saved_tensors
        p1, p2 = ctx.p
        dy2 = dy.contiguous().view(der.shape[0], -1)
        dpre = ops.gemm(dy2, p2.WT, aux=der, aux_mode=L.AUX_DACT)
        dh = ops.gemm(dpre, p1.WT)
        return dh.view(ctx.shape), dy, None, None, None


class ResidualAttentionBlock(nn.Module):
    def __init__(self, d_model: int, n_head: int, attn_mask: torch.Tensor = None):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(d_model, d_model * 4)), ("gelu", QuickGELU()),
                                              ("c_proj", nn.Linear(d_model * 4, d_model))]))
        self.ln_2 = LayerNorm(d_model)
        self.attn_mask = attn_mask
        self.n_head = n_head

    def forward(self, x: torch.Tensor):
        """x [N,B,D] (sequence first) -> [N,B,D]."""
        causal = self.attn_mask is not None  # the only mask the reference builds is the causal text mask (model.py:344-350)
        if causal:
            require_causal_mask(self.attn_mask, x.shape[0])
        if isinstance(self.attn, nn.MultiheadAttention):
            # fused path: whole block in one autograd node on batch-first memory
            spec = BlockSpec(self.ln_1, self.ln_2, _PackedInProj(self.attn), self.attn.out_proj, self.mlp.c_fc, self.mlp.c_proj,
                             self.attn.num_heads, L.ACT_QUICKGELU, causal)
            xb = x.permute(1, 0, 2)
            y = BlockFunction.apply(xb, spec, self.attn.in_proj_bias, self.attn.out_proj.bias, None, None, None, None)
            return y.permute(1, 0, 2)
        # generic path (attn replaced by PlainMultiheadAttentionLoRA, reference lora.py:93-199)
        h = _LnFunction.apply(x, self.ln_1)
        a = self.attn(h, h, h, need_weights=False, attn_mask=self.attn_mask)[0]
        x = x + a
        h2 = _LnFunction.apply(x, self.ln_2)
        return _MlpResidualFunction.apply(h2, x, self.mlp.c_fc, self.mlp.c_proj, L.ACT_QUICKGELU)


class Transformer(nn.Module):
    def __init__(self, width: int, layers: int, heads: int, attn_mask: torch.Tensor = None):
        super().__init__()
        self.width = width
        self.layers = layers
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads, attn_mask) for _ in range(layers)])

    def forward(self, x: torch.Tensor):
        return self.resblocks(x)


def _cached_cast(param, key_extra, make):
    key = (param.device, param._version, param.data_ptr(), key_extra)
    c = getattr(param, "_ngu_cache_t", None)
    if c is None or c[0] != key:
        c = (key, make())
        param._ngu_cache_t = c
    return c[1]


class _ClsHeadFunction(torch.autograd.Function):
    """ln_post(x[:, 0]) @ proj   (model.py:252-255); x is batch-first [B,N,D]."""

    @staticmethod
    def forward(ctx, x, norm, proj):
        B, N, D = x.shape
        x = x.contiguous()
        dt = x.dtype
        cls_n, mean, rstd = ops.ln_fwd(x, norm.weight.detach(), norm.bias.detach(), norm.eps, rows=B, ldx=N * D)
        Wt = _cached_cast(proj, ("T", dt), lambda: ops.cast(proj.detach().float().contiguous(), dt, transpose=True))   # [out, D]
        W = _cached_cast2(proj, dt)                                                                                   # [D, out]
        f = ops.gemm(cls_n, Wt)
        ctx.save_for_backward(x, mean, rstd)
        ctx.mods = (norm, W)
        return f

    @staticmethod
    def backward(ctx, df):
        x, mean, rstd = ctx.saved_tensors
        norm, W = ctx.mods
        B, N, D = x.shape
        dcn = ops.gemm(df.contiguous(), W)
        dx = torch.zeros_like(x)
        ops.ln_bwd(dcn, x, mean, rstd, norm.weight.detach(), rows=B, ldx=N * D, out=dx, lddx=N * D)
        return dx, None, None


def _cached_cast2(param, dt):
    key = (param.device, param._version, param.data_ptr(), dt)
    c = getattr(param, "_ngu_cache_n", None)
    if c is None or c[0] != key:
        c = (key, ops.cast(param.detach().float().contiguous(), dt))
        param._ngu_cache_n = c
    return c[1]


class VisionTransformer(nn.Module):
    def __init__(self, input_resolution: int, patch_size: int, width: int, layers: int, heads: int, output_dim: int):
        super().__init__()
        self.input_resolution = input_resolution
        self.output_dim = output_dim
        self.conv1 = nn.Conv2d(in_channels=3, out_channels=width, kernel_size=patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = LayerNorm(width)
        self.transformer = Transformer(width, layers, heads)
        self.ln_post = LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))
        self.compute_dtype = torch.bfloat16

    def embed(self, images):
        """conv1 -> cat class_embedding -> + positional_embedding -> ln_pre (model.py:233-248); frozen, no autograd graph."""
        for p in (self.conv1.weight, self.class_embedding, self.positional_embedding, self.ln_pre.weight):
            if p.requires_grad:
                raise NotImplementedError("patch embedding / ln_pre must be frozen on the ngu B200 path")
        dt = self.compute_dtype
        with torch.no_grad():
            B = images.shape[0]
            P = self.conv1.kernel_size[0]
            patches = ops.patchify(images.float().contiguous(), P, dt)
            def _w2d():   # [width, 3*P*P], K zero-padded to the patch rows' pitch (ViT-L/14: 588 -> 592)
                w = self.conv1.weight.detach().float().view(self.conv1.weight.shape[0], -1)
                return ops.cast(F.pad(w, (0, patches.shape[1] - w.shape[1])).contiguous(), dt)
            W = _cached_cast(self.conv1.weight, ("2d", dt), _w2d)
            tok = ops.gemm(patches, W)
            x = ops.assemble_tokens(tok, self.class_embedding.detach().contiguous(), self.positional_embedding.detach().contiguous(), B)
            x, _, _ = ops.ln_fwd(x, self.ln_pre.weight.detach(), self.ln_pre.bias.detach(), self.ln_pre.eps, save_stats=False)
            return x

    def forward_tokens(self, images):
        """-> batch-first [B,N,D] hidden states after the last block (blocks see the [N,B,D] view, as in the reference)."""
        x = self.embed(images)
        x = x.permute(1, 0, 2)            # NLD -> LND (a view)
        x = self.transformer(x)
        return x.permute(1, 0, 2)         # LND -> NLD (contiguous again)

    def forward(self, images):
        x = self.forward_tokens(images)
        if self.proj is None:
            raise NotImplementedError("projection-free heads are not on the reference path")
        return _ClsHeadFunction.apply(x, self.ln_post, self.proj)


class CLIP(nn.Module):
    def __init__(self, embed_dim: int, image_resolution: int, vision_layers: int, vision_width: int, vision_patch_size: int,
                 context_length: int, vocab_size: int, transformer_width: int, transformer_heads: int, transformer_layers: int):
        super().__init__()
        if isinstance(vision_layers, (tuple, list)):
            raise NotImplementedError("ModifiedResNet towers are not on the adapter fine-tuning path")
        self.context_length = context_length
        self.visual = VisionTransformer(input_resolution=image_resolution, patch_size=vision_patch_size, width=vision_width,
                                        layers=vision_layers, heads=vision_width // 64, output_dim=embed_dim)
        self.transformer = Transformer(width=transformer_width, layers=transformer_layers, heads=transformer_heads,
                                       attn_mask=self.build_attention_mask())
        self.vocab_size = vocab_size
        self.token_embedding = nn.Embedding(vocab_size, transformer_width)
        self.positional_embedding = nn.Parameter(torch.empty(self.context_length, transformer_width))
        self.ln_final = LayerNorm(transformer_width)
        self.text_projection = nn.Parameter(torch.empty(transformer_width, embed_dim))
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))
        self.initialize_parameters()

    def initialize_parameters(self):
        nn.init.normal_(self.token_embedding.weight, std=0.02)
        nn.init.normal_(self.positional_embedding, std=0.01)
        proj_std = (self.transformer.width ** -0.5) * ((2 * self.transformer.layers) ** -0.5)
        attn_std = self.transformer.width ** -0.5
        fc_std = (2 * self.transformer.width) ** -0.5
        for block in self.transformer.resblocks:
            nn.init.normal_(block.attn.in_proj_weight, std=attn_std)
            nn.init.normal_(block.attn.out_proj.weight, std=proj_std)
            nn.init.normal_(block.mlp.c_fc.weight, std=fc_std)
            nn.init.normal_(block.mlp.c_proj.weight, std=proj_std)
        nn.init.normal_(self.text_projection, std=self.transformer.width ** -0.5)

    def build_attention_mask(self):
        mask = torch.empty(self.context_length, self.context_length)
        mask.fill_(float("-inf"))
        mask.triu_(1)
        return mask

    def set_compute_dtype(self, dtype):
        self.visual.compute_dtype = dtype
        self._text_dtype = dtype
        return self

    @property
    def dtype(self):
        return self.visual.conv1.weight.dtype

    def encode_image(self, image):
        return self.visual(image)

    @torch.no_grad()
    def encode_text(self, text):
        """Frozen causal text tower, forward only (model.py:361-374): token + positional embedding, blocks with the
        causal mask, ln_final on the EOT row (argmax token id), @ text_projection."""
        if any(p.requires_grad for p in self.transformer.parameters()):
            raise NotImplementedError("the CLIP text tower is forward-only (frozen) on the ngu B200 path")
        dt = getattr(self, "_text_dtype", self.visual.compute_dtype)
        B, S = text.shape
        D = self.token_embedding.weight.shape[1]
        zero = _cached_cast(self.positional_embedding, ("zero",), lambda: torch.zeros(D, device=text.device, dtype=torch.float32))
        x = ops.embed_tokens(text.long().contiguous(), self.token_embedding.weight.detach(), self.positional_embedding.detach(), zero, dt)
        x = x.view(B, S, D).permute(1, 0, 2)
        x = self.transformer(x).permute(1, 0, 2).contiguous()
        eot = x[torch.arange(B, device=x.device), text.argmax(dim=-1)].contiguous()
        eot, _, _ = ops.ln_fwd(eot, self.ln_final.weight.detach(), self.ln_final.bias.detach(), self.ln_final.eps, save_stats=False)
        Wt = _cached_cast(self.text_projection, ("T", dt),
                          lambda: ops.cast(self.text_projection.detach().float().contiguous(), dt, transpose=True))
        return ops.gemm(eot, Wt)

    def forward(self, image, text):
        fi, ft = self.encode_image(image), self.encode_text(text)
        fi = fi / fi.norm(dim=1, keepdim=True)
        ft = ft / ft.norm(dim=1, keepdim=True)
        logits = self.logit_scale.exp() * fi.float() @ ft.float().t()
        return logits, logits.t()
