"""Functional CPU restatement of the hot path, driven by a state dict with open_clip-style keys.

Every function cites the reference lines it follows.  Nothing here is used by the product path.
"""
import math

import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------------
# Mona  (reference: src/adapters/mona.py)
# ---------------------------------------------------------------------------------------------------
def mona_conv_stage(h, p, prefix, hw, has_cls):
    """h [B,N,C] -> conv stage output [B,N,C].  mona.py:85-93 (BaselineMonaOp.forward) applied to the
    spatial tokens only when a CLS token is present (mona.py:129-139), to all tokens otherwise (:140-144)."""
    B, N, C = h.shape
    H, W = hw
    sp = h[:, 1:, :] if has_cls else h
    z = sp.reshape(B, H, W, C).permute(0, 3, 1, 2)                                  # NCHW, mona.py:135
    ident = z
    ac = f"{prefix}adapter_conv."
    if f"{ac}freq_filter" in p:                                                     # mona.py:283-286 / :395-398
        zf = torch.fft.rfft2(z, dim=(-2, -1)) * p[f"{ac}freq_filter"].view(1, -1, 1, 1)
        z = torch.fft.irfft2(zf, s=(H, W), dim=(-2, -1))
    branches = [F.conv2d(z, p[f"{ac}{name}.weight"], p[f"{ac}{name}.bias"], padding=pad, groups=C)
                for name, pad in (("conv1", 1), ("conv2", 2), ("conv3", 3))]        # 3x3, 5x5, 7x7 depthwise
    if f"{ac}noise_estimator.1.weight" in p:                                        # mona.py:170-176, :187-192
        g = z.mean(dim=(2, 3), keepdim=True)
        g = F.relu(F.conv2d(g, p[f"{ac}noise_estimator.1.weight"], p[f"{ac}noise_estimator.1.bias"]))
        wts = torch.softmax(F.conv2d(g, p[f"{ac}noise_estimator.3.weight"], p[f"{ac}noise_estimator.3.bias"]), dim=1)
        z = sum(b * wts[:, i:i + 1] for i, b in enumerate(branches)) + ident
    else:
        z = sum(branches) / 3.0 + ident                                             # mona.py:89
    z = z + F.conv2d(z, p[f"{ac}projector.weight"], p[f"{ac}projector.bias"])       # mona.py:91-93
    sp = z.permute(0, 2, 3, 1).reshape(B, H * W, C)
    return torch.cat([h[:, :1, :], sp], 1) if has_cls else sp


def mona(x, p, prefix, hw, has_cls=True):
    """Batch-first Mona: x [B,N,D] -> [B,N,D].  mona.py:115-151 with the wrapper permutes (:54-67) cancelled.
    Eval-mode (dropout off, mona.py:147)."""
    D = x.shape[-1]
    u = F.layer_norm(x, (D,), p[f"{prefix}norm.weight"], p[f"{prefix}norm.bias"], 1e-5) * p[f"{prefix}gamma"] \
        + x * p[f"{prefix}gammax"]                                                  # mona.py:125
    h = F.linear(u, p[f"{prefix}project1.weight"], p[f"{prefix}project1.bias"])     # mona.py:127
    h = mona_conv_stage(h, p, prefix, hw, has_cls)
    h = F.gelu(h)                                                                   # mona.py:146 (exact erf)
    return x + F.linear(h, p[f"{prefix}project2.weight"], p[f"{prefix}project2.bias"])  # mona.py:148-150


# ---------------------------------------------------------------------------------------------------
# LoRA  (reference: src/adapters/lora.py)
# ---------------------------------------------------------------------------------------------------
def lora_linear(x, p, prefix, r, alpha):
    """y = x W^T + b + (alpha/sqrt r) * x (B A)^T.  lora.py:78-90 with scaling from lora.py:19-20; eval mode."""
    y = F.linear(x, p[f"{prefix}weight"], p.get(f"{prefix}bias"))
    if r > 0 and f"{prefix}w_lora_A" in p:
        s = alpha / math.sqrt(r)
        y = y + (x @ p[f"{prefix}w_lora_A"].t()) @ p[f"{prefix}w_lora_B"].t() * s
    return y


# ---------------------------------------------------------------------------------------------------
# timm ViT block / tower  (pinned dep timm 1.0.20 `vit_base_patch16_224`; restated)
# ---------------------------------------------------------------------------------------------------
def vit_block(x, p, prefix, heads, lora=None, eps=1e-6):
    """Pre-LN block: x += proj(SDPA(qkv(LN1 x))); x += fc2(gelu(fc1(LN2 x))).  [pinned-dep knowledge; the trunk is checked
    against transformers.ViTModel in tests/test_cpu_oracle.py]"""
    B, N, D = x.shape
    dh = D // heads
    r, alpha = lora if lora else (0, 1)
    xn = F.layer_norm(x, (D,), p[f"{prefix}norm1.weight"], p[f"{prefix}norm1.bias"], eps)
    qkv = lora_linear(xn, p, f"{prefix}attn.qkv.", r, alpha).reshape(B, N, 3, heads, dh).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    att = torch.softmax((q @ k.transpose(-1, -2)) * dh ** -0.5, -1) @ v             # == F.scaled_dot_product_attention
    att = att.transpose(1, 2).reshape(B, N, D)
    x = x + lora_linear(att, p, f"{prefix}attn.proj.", r, alpha)
    xn = F.layer_norm(x, (D,), p[f"{prefix}norm2.weight"], p[f"{prefix}norm2.bias"], eps)
    h = F.gelu(F.linear(xn, p[f"{prefix}mlp.fc1.weight"], p[f"{prefix}mlp.fc1.bias"]))
    return x + F.linear(h, p[f"{prefix}mlp.fc2.weight"], p[f"{prefix}mlp.fc2.bias"])


def encode_image(p, images, cfg, taps=None, adapters=None):
    """open_clip TimmModel: patch_embed -> cat cls -> +pos -> blocks (each followed by Mona when injected,
    mona.py:667-676) -> final norm -> CLS pool -> head.proj.  [pinned-dep knowledge for the trunk]
    `adapters`: optional list of callables (x [B,N,D], hw) -> [B,N,D], one per block, used INSTEAD of the restated Mona:
    bench.py's reference arm passes the reference's own BatchFirstMonaWrapper modules here."""
    t = "visual.trunk."
    P = cfg["patch"]
    x = F.conv2d(images, p[f"{t}patch_embed.proj.weight"], p[f"{t}patch_embed.proj.bias"], stride=P)
    B, D, gh, gw = x.shape
    x = x.flatten(2).transpose(1, 2)
    x = torch.cat([p[f"{t}cls_token"].expand(B, -1, -1), x], 1) + p[f"{t}pos_embed"]
    for i in range(cfg["depth"]):
        x = vit_block(x, p, f"{t}blocks.{i}.", cfg["heads"], cfg.get("lora"))
        mp = f"{t}blocks.{i}.mona.clip_mona."
        if adapters is not None:
            x = adapters[i](x, (gh, gw))
        elif f"{mp}gamma" in p:
            x = mona(x, p, mp, (gh, gw), True)
        if taps is not None:
            taps.append(x)
    x = F.layer_norm(x, (D,), p[f"{t}norm.weight"], p[f"{t}norm.bias"], 1e-6)
    return F.linear(x[:, 0], p["visual.head.proj.weight"])


# ---------------------------------------------------------------------------------------------------
# OpenAI CLIP towers (reference: src/third_party/openai_clip/model.py, vendored in the reference tree)
# ---------------------------------------------------------------------------------------------------
def clip_block(x, p, prefix, heads, causal=False, lora=None):
    """ResidualAttentionBlock on batch-first x [B,N,D] (the reference runs it sequence-first; the arithmetic is
    per token/per image so the layouts are equivalent).  model.py:199-202: x += attn(ln_1 x); x += mlp(ln_2 x) with
    nn.MultiheadAttention (packed in_proj q|k|v, model.py:195-197), QuickGELU x*sigmoid(1.702x) (:172-174), LN eps 1e-5.
    With `lora` = (r, alpha) the attention is PlainMultiheadAttentionLoRA (lora.py:155-199): separate q/k/v/proj
    LinearLoRA projections."""
    B, N, D = x.shape
    dh = D // heads
    h = F.layer_norm(x, (D,), p[f"{prefix}ln_1.weight"], p[f"{prefix}ln_1.bias"], 1e-5)
    if lora is None:
        qkv = F.linear(h, p[f"{prefix}attn.in_proj_weight"], p[f"{prefix}attn.in_proj_bias"])
        q, k, v = qkv.split(D, dim=-1)
    else:
        r, alpha = lora
        q = lora_linear(h, p, f"{prefix}attn.q_proj.", r, alpha)
        k = lora_linear(h, p, f"{prefix}attn.k_proj.", r, alpha)
        v = lora_linear(h, p, f"{prefix}attn.v_proj.", r, alpha)
    sh = lambda t: t.reshape(B, N, heads, dh).transpose(1, 2)
    s = (sh(q) @ sh(k).transpose(-1, -2)) * dh ** -0.5
    if causal:
        s = s + torch.full((N, N), float("-inf"), dtype=s.dtype).triu(1)
    a = (torch.softmax(s, -1) @ sh(v)).transpose(1, 2).reshape(B, N, D)
    if lora is None:
        a = F.linear(a, p[f"{prefix}attn.out_proj.weight"], p[f"{prefix}attn.out_proj.bias"])
    else:
        a = lora_linear(a, p, f"{prefix}attn.proj.", lora[0], lora[1])
    x = x + a
    h = F.layer_norm(x, (D,), p[f"{prefix}ln_2.weight"], p[f"{prefix}ln_2.bias"], 1e-5)
    h = F.linear(h, p[f"{prefix}mlp.c_fc.weight"], p[f"{prefix}mlp.c_fc.bias"])
    h = h * torch.sigmoid(1.702 * h)
    return x + F.linear(h, p[f"{prefix}mlp.c_proj.weight"], p[f"{prefix}mlp.c_proj.bias"])


def clip_encode_image(p, images, cfg, taps=None, tap_layers=()):
    """VisionTransformer.forward, model.py:233-257, with Mona after each block when injected (mona.py:563-571:
    adapter gets [N,B,D] and (grid, grid))."""
    v = "visual."
    P = cfg["patch"]
    x = F.conv2d(images, p[f"{v}conv1.weight"], None, stride=P)
    B, D, gh, gw = x.shape
    x = x.reshape(B, D, -1).permute(0, 2, 1)
    x = torch.cat([p[f"{v}class_embedding"].expand(B, 1, D), x], 1) + p[f"{v}positional_embedding"]
    x = F.layer_norm(x, (D,), p[f"{v}ln_pre.weight"], p[f"{v}ln_pre.bias"], 1e-5)
    for i in range(cfg["depth"]):
        x = clip_block(x, p, f"{v}transformer.resblocks.{i}.", cfg["heads"], False, cfg.get("lora"))
        mp = f"{v}transformer.resblocks.{i}.mona."
        if f"{mp}gamma" in p:
            x = mona(x, p, mp, (gh, gw), True)
        if taps is not None and i in tap_layers:     # clipseg_adapter.py:60-68 (hidden states after listed blocks, NLD)
            taps.append(x)
    x = F.layer_norm(x[:, 0], (D,), p[f"{v}ln_post.weight"], p[f"{v}ln_post.bias"], 1e-5)
    return x @ p[f"{v}proj"]


def clip_encode_text(p, text, cfg):
    """CLIP.encode_text, model.py:361-374 (causal mask from build_attention_mask :344-350, EOT = argmax token id)."""
    x = p["token_embedding.weight"][text] + p["positional_embedding"]
    D = x.shape[-1]
    for i in range(cfg["text_layers"]):
        x = clip_block(x, p, f"transformer.resblocks.{i}.", cfg["text_heads"], True)
    x = F.layer_norm(x, (D,), p["ln_final.weight"], p["ln_final.bias"], 1e-5)
    return x[torch.arange(x.shape[0]), text.argmax(-1)] @ p["text_projection"]


# ---------------------------------------------------------------------------------------------------
# BERT text tower (pinned deps transformers 4.57.1 BertModel + open_clip HFTextEncoder; restated)
# ---------------------------------------------------------------------------------------------------
def encode_text(p, ids, cfg, pad_token_id=0):
    """BertModel (post-LN, eps 1e-12, exact GELU, absolute positions, token_type 0; keys at pad positions are masked,
    open_clip HFTextEncoder.forward `attn_mask = (x != pad_token_id)`) -> CLS last-hidden-state pooler -> MLP proj
    (768->640->GELU->512, no bias).  Eval mode.  [pinned-dep knowledge; the encoder stack incl. the mask is checked against
    transformers.BertModel in tests/test_cpu_oracle.py]"""
    t = "text.transformer."
    B, S = ids.shape
    heads = cfg["text_heads"]
    x = p[f"{t}embeddings.word_embeddings.weight"][ids] + p[f"{t}embeddings.position_embeddings.weight"][:S] \
        + p[f"{t}embeddings.token_type_embeddings.weight"][0]
    D = x.shape[-1]
    dh = D // heads
    x = F.layer_norm(x, (D,), p[f"{t}embeddings.LayerNorm.weight"], p[f"{t}embeddings.LayerNorm.bias"], 1e-12)
    for i in range(cfg["text_layers"]):
        l = f"{t}encoder.layer.{i}."
        tr_, ta_ = cfg.get("text_lora") or (0, 1)      # LoRA on the BERT projections (lora.py:317-367, tune_text_encoder)
        def heads_(name):
            return lora_linear(x, p, f"{l}attention.self.{name}.", tr_, ta_).reshape(B, S, heads, dh).transpose(1, 2)
        q, k, v = heads_("query"), heads_("key"), heads_("value")
        sc = q @ k.transpose(-1, -2) / math.sqrt(dh)
        sc = sc.masked_fill((ids == pad_token_id)[:, None, None, :], float("-inf"))
        a = (torch.softmax(sc, -1) @ v).transpose(1, 2).reshape(B, S, D)
        a = lora_linear(a, p, f"{l}attention.output.dense.", tr_, ta_)
        x = F.layer_norm(x + a, (D,), p[f"{l}attention.output.LayerNorm.weight"], p[f"{l}attention.output.LayerNorm.bias"], 1e-12)
        h = F.gelu(F.linear(x, p[f"{l}intermediate.dense.weight"], p[f"{l}intermediate.dense.bias"]))
        h = F.linear(h, p[f"{l}output.dense.weight"], p[f"{l}output.dense.bias"])
        x = F.layer_norm(x + h, (D,), p[f"{l}output.LayerNorm.weight"], p[f"{l}output.LayerNorm.bias"], 1e-12)
    c = x[:, 0]
    return F.linear(F.gelu(F.linear(c, p["text.proj.0.weight"])), p["text.proj.2.weight"])


# ---------------------------------------------------------------------------------------------------
# InfoNCE  (reference: src/losses/losses.py:23-47)
# ---------------------------------------------------------------------------------------------------
def info_nce(img, txt, temperature=0.07):
    i = img / img.norm(dim=1, keepdim=True).clamp_min(1e-12)       # F.normalize, losses.py:25-26
    t = txt / txt.norm(dim=1, keepdim=True).clamp_min(1e-12)
    logits = i @ t.t() / temperature                               # losses.py:34
    lab = torch.arange(img.shape[0])                               # losses.py:38
    li = (torch.logsumexp(logits, 1) - logits[lab, lab]).mean()    # CE(logits, arange), losses.py:41
    lt = (torch.logsumexp(logits, 0) - logits[lab, lab]).mean()    # CE(logits^T, arange), losses.py:42
    return (li + lt) / 2, logits                                   # losses.py:45


def zero_shot_predict(img_feat, class_text_feats):
    """Prompt-ensemble zero-shot: mean over prompts of 100 * Ihat . That^T per class, argmax over classes
    (src/models/biomedclip/zero_shot.py:176-228).  class_text_feats: list of [n_prompts, E] per class."""
    i = img_feat / img_feat.norm(dim=-1, keepdim=True)
    scores = []
    for tf in class_text_feats:
        t = tf / tf.norm(dim=-1, keepdim=True)
        scores.append((100.0 * i @ t.t()).mean(dim=1))
    return torch.stack(scores, 1).argmax(1)


# ---------------------------------------------------------------------------------------------------
def training_loss(p, images, ids, cfg, temperature=0.07):
    """One micro-step of src/models/biomedclip/finetune.py:272-279 (eval-mode numerics)."""
    fi = encode_image(p, images, cfg)
    with torch.no_grad():
        ft = encode_text(p, ids, cfg)
    loss, logits = info_nce(fi, ft, temperature)
    return loss, fi, ft, logits


def loss_and_grads(state_dict, images, ids, cfg, trainable, dtype=torch.float64, temperature=0.07):
    p = {k: v.detach().to(dtype).clone() if v.is_floating_point() else v.detach().clone() for k, v in state_dict.items()}
    for k in trainable:
        p[k].requires_grad_(True)
    loss, fi, ft, logits = training_loss(p, images.to(dtype), ids, cfg, temperature)
    grads = torch.autograd.grad(loss, [p[k] for k in trainable])
    return loss.detach(), fi.detach(), ft.detach(), logits.detach(), dict(zip(trainable, [g.detach() for g in grads]))
