"""CPU: the oracle against the golden vectors produced from the reference's own modules
(oracle/make_golden.py), plus identities from SURVEY.md §8c."""
import math

import pytest
import torch
import torch.nn.functional as F

from conftest import relerr
from oracle import functional as OF


@pytest.mark.parametrize("name", ["mona_cls", "mona_nocls", "mona_noise", "mona_freq", "mona_hybrid"])
def test_oracle_mona_matches_reference_golden(golden, name):
    g = golden(name)
    p = {k: v.double().requires_grad_(True) for k, v in g["state"].items()}
    x = g["x"].double().requires_grad_(True)
    y = OF.mona(x, p, "clip_mona.", g["hw"], g["has_cls"])
    assert relerr(y, g["y"]) < 1e-6
    keys = list(g["grads"].keys())
    grads = torch.autograd.grad((y * g["gy"].double()).sum(), [x] + [p[k] for k in keys])
    assert relerr(grads[0], g["dx"]) < 1e-5
    for k, gr in zip(keys, grads[1:]):
        assert relerr(gr, g["grads"][k]) < 1e-5, k


def test_oracle_lora_matches_reference_golden(golden):
    g = golden("lora_linear")
    p = {f"l.{k}": v.double() for k, v in g["state"].items()}
    y = OF.lora_linear(g["x"].double(), p, "l.", g["r"], g["alpha"])
    assert relerr(y, g["y"]) < 1e-6
    assert abs(g["alpha"] / math.sqrt(g["r"]) - 11.3137085) < 1e-6  # s = 32/sqrt(8)


@pytest.mark.parametrize("name", ["infonce_b8", "infonce_b37"])
def test_oracle_infonce_matches_reference_golden(golden, name):
    g = golden(name)
    I, T = g["I"].double().requires_grad_(True), g["T"].double().requires_grad_(True)
    loss, _ = OF.info_nce(I, T, g["temperature"])
    gI, gT = torch.autograd.grad(loss, [I, T])
    assert abs(float(loss) - float(g["loss"])) < 1e-6
    assert relerr(gI, g["dI"]) < 1e-5 and relerr(gT, g["dT"]) < 1e-5


def test_merged_stencil_identity():
    """(dw3+dw5+dw7)/3 + x == one 7x7 depthwise with K=(pad(k3)+pad(k5)+k7)/3+delta — the kernel's stencil."""
    torch.manual_seed(0)
    C = 8
    x = torch.randn(2, C, 14, 14, dtype=torch.float64)
    k3, k5, k7 = (torch.randn(C, 1, k, k, dtype=torch.float64) for k in (3, 5, 7))
    b3, b5, b7 = (torch.randn(C, dtype=torch.float64) for _ in range(3))
    ref = (F.conv2d(x, k3, b3, padding=1, groups=C) + F.conv2d(x, k5, b5, padding=2, groups=C) + F.conv2d(x, k7, b7, padding=3, groups=C)) / 3 + x
    K = (F.pad(k3, (2, 2, 2, 2)) + F.pad(k5, (1, 1, 1, 1)) + k7) / 3
    K[:, 0, 3, 3] += 1
    got = F.conv2d(x, K, (b3 + b5 + b7) / 3, padding=3, groups=C)
    assert relerr(got, ref) < 1e-12


def test_lora_is_identity_at_init_and_mona_is_not():
    from src.adapters import LinearLoRA, BaselineMona
    torch.manual_seed(0)
    lin = torch.nn.Linear(32, 48)
    ll = LinearLoRA(lin, r=8, lora_alpha=32, dropout_rate=0.1)
    assert torch.count_nonzero(ll.w_lora_B) == 0 and not ll.weight.requires_grad and ll.bias.requires_grad
    assert torch.equal(ll.weight, lin.weight)
    m = BaselineMona(64, 16)
    assert float(m.gamma[0]) == pytest.approx(1e-6) and float(m.gammax[0]) == 1.0
    assert len(list(m.parameters())) == 16


def test_oracle_global_batch_equals_concat():
    """DP semantics: loss on the concatenated global features == what every rank computes."""
    torch.manual_seed(0)
    I, T = torch.randn(8, 16, dtype=torch.float64), torch.randn(8, 16, dtype=torch.float64)
    full, _ = OF.info_nce(I, T)
    halves = [OF.info_nce(I[:4], T[:4])[0], OF.info_nce(I[4:], T[4:])[0]]
    assert abs(float(full) - float(sum(halves) / 2)) > 1e-3  # local-only losses differ: global negatives matter


@pytest.mark.parametrize("tag", ["clip_mona", "clip_lora"])
def test_oracle_clip_matches_vendored_reference_golden(golden, tag):
    """oracle restatement of the vendored OpenAI CLIP towers vs outputs of the reference model itself."""
    g = golden(tag)
    p = {k: (v.double().requires_grad_(k in g["trainable"]) if v.is_floating_point() else v) for k, v in g["state"].items()}
    fi = OF.clip_encode_image(p, g["images"].double(), g["cfg"])
    ft = OF.clip_encode_text(p, g["text"], g["cfg"])
    assert relerr(fi, g["fi"]) < 1e-5 and relerr(ft, g["ft"]) < 1e-5
    grads = torch.autograd.grad((fi * g["gi"].double()).sum(), [p[n] for n in g["trainable"]])
    for n, gr in zip(g["trainable"], grads):
        assert relerr(gr, g["grads"][n]) < 1e-4, n


def test_oracle_bert_stack_with_padding_matches_transformers():
    """The oracle's BERT restatement (encode_text: embeddings, post-LN encoder, exact GELU, key-padding mask, CLS pooling)
    against transformers.BertModel with random weights on a right-padded batch -- pins the [pinned-dep knowledge] part
    of the text tower, including the attention mask open_clip's HFTextEncoder builds from pad_token_id."""
    transformers = pytest.importorskip("transformers")
    import oracle.functional as OF
    import torch.nn.functional as F
    torch.manual_seed(31)
    cfg = transformers.BertConfig(vocab_size=100, hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128,
                                  max_position_embeddings=32, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, pad_token_id=0)
    bert = transformers.BertModel(cfg, add_pooling_layer=False).eval().double()
    ids = torch.randint(5, 100, (4, 12))
    ids[1, 7:] = 0
    ids[2, 3:] = 0
    ids[3, 11:] = 0
    with torch.no_grad():
        hid = bert(input_ids=ids, attention_mask=(ids != 0).long()).last_hidden_state
    W0, W2 = torch.randn(40, 64, dtype=torch.float64) * 0.1, torch.randn(16, 40, dtype=torch.float64) * 0.1
    ref = F.linear(F.gelu(F.linear(hid[:, 0], W0)), W2)
    p = {"text.transformer." + k: v for k, v in bert.state_dict().items()}
    p["text.proj.0.weight"], p["text.proj.2.weight"] = W0, W2
    got = OF.encode_text(p, ids, dict(text_layers=2, text_heads=4))
    assert float((got - ref).abs().max()) < 1e-10
    # without the mask the padded rows differ (the check above is not vacuous)
    nomask = OF.encode_text(p, ids, dict(text_layers=2, text_heads=4), pad_token_id=-1)
    assert float((nomask[1] - ref[1]).abs().max()) > 1e-6


def test_oracle_vit_trunk_matches_transformers_vit():
    """The oracle's ViT trunk (encode_image without adapters: conv patch embedding, CLS token + learned positions, pre-LN
    blocks with fused qkv, exact GELU, final LayerNorm, CLS pooling, bias-free head) against transformers.ViTModel -- an
    independent implementation of the architecture timm's `vit_base_patch16_224` implements (timm itself is not
    installed here).  Random weights mapped name by name, fp64."""
    transformers = pytest.importorskip("transformers")
    import oracle.functional as OF
    import torch.nn.functional as F
    torch.manual_seed(32)
    D, depth, heads, res, P = 64, 2, 4, 32, 16
    cfg = transformers.ViTConfig(hidden_size=D, num_hidden_layers=depth, num_attention_heads=heads, intermediate_size=4 * D,
                                 image_size=res, patch_size=P, layer_norm_eps=1e-6, hidden_act="gelu", qkv_bias=True,
                                 hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    vit = transformers.ViTModel(cfg, add_pooling_layer=False).eval().double()
    with torch.no_grad():
        for p_ in vit.parameters():
            p_.copy_(torch.randn_like(p_) * 0.1)
    hf = vit.state_dict()
    t = "visual.trunk."
    p = {f"{t}patch_embed.proj.weight": hf["embeddings.patch_embeddings.projection.weight"],
         f"{t}patch_embed.proj.bias": hf["embeddings.patch_embeddings.projection.bias"],
         f"{t}cls_token": hf["embeddings.cls_token"], f"{t}pos_embed": hf["embeddings.position_embeddings"],
         f"{t}norm.weight": hf["layernorm.weight"], f"{t}norm.bias": hf["layernorm.bias"]}
    for i in range(depth):
        h, o = f"encoder.layer.{i}.", f"{t}blocks.{i}."
        for w in ("weight", "bias"):
            p[f"{o}norm1.{w}"] = hf[f"{h}layernorm_before.{w}"]
            p[f"{o}norm2.{w}"] = hf[f"{h}layernorm_after.{w}"]
            p[f"{o}attn.qkv.{w}"] = torch.cat([hf[f"{h}attention.attention.{n}.{w}"] for n in ("query", "key", "value")], 0)
            p[f"{o}attn.proj.{w}"] = hf[f"{h}attention.output.dense.{w}"]
            p[f"{o}mlp.fc1.{w}"] = hf[f"{h}intermediate.dense.{w}"]
            p[f"{o}mlp.fc2.{w}"] = hf[f"{h}output.dense.{w}"]
    W = torch.randn(16, D, dtype=torch.float64) * 0.1
    p["visual.head.proj.weight"] = W
    images = torch.rand(3, 3, res, res, dtype=torch.float64)
    with torch.no_grad():
        ref = F.linear(vit(pixel_values=images).last_hidden_state[:, 0], W)
    got = OF.encode_image(p, images, dict(patch=P, depth=depth, heads=heads))
    assert float((got - ref).abs().max()) < 1e-10
