"""timm-compatible ViT tower whose transformer block runs on the B200 kernels.

The reference gets this arithmetic from un-vendored pinned dependencies (timm 1.0.20
`vit_base_patch16_224` inside open-clip-torch 3.2.0's TimmModel).  The attribute
names below are exactly the ones the reference dereferences when it injects adapters
(src/adapters/mona.py:620-630, src/adapters/lora.py:284-313): trunk.embed_dim,
trunk.patch_embed.num_patches, trunk.cls_token, trunk.pos_embed, trunk.blocks[i].{norm1, attn.qkv,
attn.proj, norm2, mlp.fc1, mlp.fc2}, trunk.norm — so `inject_mona_variant_to_open_clip` and
`inject_lora_to_biomedclip` work on this tower unchanged.

Block semantics (pre-LN, eps 1e-6, exact GELU, qkv bias, no LayerScale/DropPath):
    x = x + proj(SDPA(split(qkv(norm1 x))));  x = x + fc2(gelu(fc1(norm2 x)))
Forward = 2 LN kernels + 4 tcgen05 GEMMs (bias / GELU(+pre-act save) / residual fused in the epilogue)
+ 1 attention kernel.  Backward is hand written: frozen weights => dgrad only; the residual-stream
adds are fused into the LN-backward kernels; dGELU is fused into the fc2-dgrad epilogue.
"""
import math

import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .linear import Proj, proj_fwd, proj_bwd, frozen_copies


def _seed(training):
    return int(torch.randint(0, 2 ** 62, (1,)).item()) if training else 0


class BlockSpec:
    """What BlockFunction needs from a transformer block, independent of the attribute names of the host module:
    timm Block (norm1/attn.qkv/attn.proj/norm2/mlp.fc1/mlp.fc2) and the OpenAI-CLIP ResidualAttentionBlock
    (ln_1/attn.in_proj_*/attn.out_proj/ln_2/mlp.c_fc/mlp.c_proj) both map onto it."""

    __slots__ = ("ln1", "ln2", "qkv", "proj", "fc1", "fc2", "heads", "act", "causal")

    def __init__(self, ln1, ln2, qkv, proj, fc1, fc2, heads, act, causal=False):
        self.ln1, self.ln2, self.qkv, self.proj, self.fc1, self.fc2 = ln1, ln2, qkv, proj, fc1, fc2
        self.heads, self.act, self.causal = heads, act, causal


class BlockFunction(torch.autograd.Function):
    """One pre-LN transformer block on [B,N,D] with frozen base weights (+ optional LoRA on qkv / proj)."""

    @staticmethod
    def forward(ctx, x, spec, qkv_bias, proj_bias, qA, qB, pA, pB):
        B, N, D = x.shape
        H = spec.heads
        dh = D // H
        M = B * N
        x2 = x.contiguous().view(M, D)
        dt = x2.dtype
        pq, pp = Proj(spec.qkv, dt), Proj(spec.proj, dt)
        p1, p2 = Proj(spec.fc1, dt), Proj(spec.fc2, dt)
        act = spec.act
        sq, sp = _seed(pq.p > 0), _seed(pp.p > 0)

        xn, m1, r1 = ops.ln_fwd(x2, spec.ln1.weight.detach(), spec.ln1.bias.detach(), spec.ln1.eps)
        qkv, sv_q = proj_fwd(xn, pq, seed=sq)
        ao, lse = ops.attn_fwd_packed(qkv, B, N, H, dh, causal=spec.causal)
        x1, sv_p = proj_fwd(ao, pp, aux=x2, aux_mode=L.AUX_RESIDUAL, seed=sp)
        xn2, m2, r2 = ops.ln_fwd(x1, spec.ln2.weight.detach(), spec.ln2.bias.detach(), spec.ln2.eps)
        (hact, hder), _ = proj_fwd(xn2, p1, act=act, save_pre=True)   # hder = act'(pre), what backward multiplies by
        y, _ = proj_fwd(hact, p2, aux=x1, aux_mode=L.AUX_RESIDUAL)

        ctx.save_for_backward(x2, m1, r1, qkv, ao, lse, x1, m2, r2, hder)
        ctx.lora_saved = (sv_q, sv_p)
        ctx.projs = (pq, pp, p1, p2)
        ctx.spec = spec
        ctx.dims = (B, N, D, H, dh)
        return y.view(B, N, D)

    @staticmethod
    def backward(ctx, dy):
        x2, m1, r1, qkv, ao, lse, x1, m2, r2, hder = ctx.saved_tensors
        pq, pp, p1, p2 = ctx.projs
        sv_q, sv_p = ctx.lora_saved
        spec = ctx.spec
        B, N, D, H, dh = ctx.dims
        need = ctx.needs_input_grad
        dy2 = dy.contiguous().view(B * N, D)
        # MLP:  y = x1 + fc2(act(fc1(LN2 x1)))
        dhpre = ops.gemm(dy2, p2.WT, aux=hder, aux_mode=L.AUX_DACT)                   # (dy W2) * act'(pre)
        dxn2 = ops.gemm(dhpre, p1.WT)
        dx1 = ops.ln_bwd(dxn2, x1, m2, r2, spec.ln2.weight.detach(), dres=dy2)        # dy + LN2bwd
        # attention:  x1 = x + proj(attn(qkv(LN1 x)))
        dao, dpb, dpA, dpB = proj_bwd(dx1, pp, sv_p, need_dx=True, need_bias=need[3])
        dqkv = ops.attn_bwd_packed(qkv, ao, lse, dao, B, N, H, dh, causal=spec.causal)
        dxn, dqb, dqA, dqB = proj_bwd(dqkv, pq, sv_q, need_dx=need[0], need_bias=need[2])
        dx = None
        if need[0]:
            dx = ops.ln_bwd(dxn, x2, m1, r1, spec.ln1.weight.detach(), dres=dx1).view(B, N, D)
        return dx, None, dqb, dpb, dqA, dqB, dpA, dpB


class Attention(nn.Module):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)


class Block(nn.Module):
    """Call convention: block(x [B,N,D], **kwargs) -> [B,N,D] (what the Mona injection wraps)."""

    act_kind = L.ACT_GELU

    def __init__(self, dim, num_heads, mlp_ratio=4.0, eps=1e-6):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=eps)
        self.attn = Attention(dim, num_heads)
        self.norm2 = nn.LayerNorm(dim, eps=eps)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))

    def forward(self, x, **kwargs):
        q, p = self.attn.qkv, self.attn.proj
        spec = BlockSpec(self.norm1, self.norm2, q, p, self.mlp.fc1, self.mlp.fc2, self.attn.num_heads, self.act_kind)
        return BlockFunction.apply(x, spec, q.bias, p.bias,
                                   getattr(q, "w_lora_A", None), getattr(q, "w_lora_B", None),
                                   getattr(p, "w_lora_A", None), getattr(p, "w_lora_B", None))


class PatchEmbed(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        self.img_size = (img_size, img_size)
        self.patch_size = (patch_size, patch_size)
        self.grid_size = (img_size // patch_size, img_size // patch_size)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)


class HeadFunction(torch.autograd.Function):
    """features = LN(x[:, 0]) @ Wproj^T — final norm applied to the CLS rows only (CLS pooling), then the
    bias-free 768->512 head (open_clip TimmModel head.proj)."""

    @staticmethod
    def forward(ctx, x, norm, proj):
        B, N, D = x.shape
        x = x.contiguous()
        dt = x.dtype
        cls_n, mean, rstd = ops.ln_fwd(x, norm.weight.detach(), norm.bias.detach(), norm.eps, rows=B, ldx=N * D)
        W, WT = frozen_copies(proj.weight, dt)
        f = ops.gemm(cls_n, W, bias=(proj.bias.detach() if proj.bias is not None else None))
        ctx.save_for_backward(x, mean, rstd)
        ctx.mods = (norm, WT)
        return f

    @staticmethod
    def backward(ctx, df):
        x, mean, rstd = ctx.saved_tensors
        norm, WT = ctx.mods
        B, N, D = x.shape
        dcn = ops.gemm(df.contiguous(), WT)
        dx = torch.zeros_like(x)
        ops.ln_bwd(dcn, x, mean, rstd, norm.weight.detach(), rows=B, ldx=N * D, out=dx, lddx=N * D)
        return dx, None, None


class VisionTransformer(nn.Module):
    """timm `vit_base_patch16_224`-shaped trunk (class token, learned pos-embed, CLS pooling, final norm)."""

    def __init__(self, img_size=224, patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0):
        super().__init__()
        self.embed_dim = self.num_features = embed_dim
        self.patch_embed = PatchEmbed(img_size, patch_size, 3, embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches + 1, embed_dim))
        self.pos_drop = nn.Dropout(p=0.0)
        self.blocks = nn.Sequential(*[Block(embed_dim, num_heads, mlp_ratio) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        self.compute_dtype = torch.bfloat16

    def embed(self, images):
        """images fp32 [B,3,R,R] -> tokens [B, 1+np, D]; frozen, input needs no grad => no autograd graph."""
        pe = self.patch_embed
        if pe.proj.weight.requires_grad or self.pos_embed.requires_grad or self.cls_token.requires_grad:
            raise NotImplementedError("patch embedding / position embedding must be frozen on the ngu B200 path")
        dt = self.compute_dtype
        with torch.no_grad():
            B = images.shape[0]
            patches = ops.patchify(images.float().contiguous(), pe.patch_size[0], dt)
            w2d = pe.proj.weight.view(pe.proj.weight.shape[0], -1)
            W = _frozen_2d(pe.proj.weight, w2d, dt)
            tok = ops.gemm(patches, W, bias=pe.proj.bias.detach())
            return ops.assemble_tokens(tok, self.cls_token.detach().view(-1).contiguous(),
                                       self.pos_embed.detach().view(-1, self.embed_dim).contiguous(), B)

    def forward_features(self, images):
        return self.blocks(self.embed(images))

    def forward(self, images):
        return self.forward_features(images)


def _frozen_2d(param, view2d, dtype):
    key = (dtype, param.device, param._version, param.data_ptr())
    cache = getattr(param, "_ngu_cache2d", None)
    if cache is None or cache[0] != key:
        cache = (key, ops.cast(view2d.detach().float().contiguous(), dtype))
        param._ngu_cache2d = cache
    return cache[1]
