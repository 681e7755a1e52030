"""GPU parity of the drop-in modules (Mona, LoRA, block, full model) against golden vectors made from
the reference's own modules and against the CPU oracle."""
import pytest
import torch

from conftest import relerr, TOL, GTOL

pytestmark = pytest.mark.gpu
DTYPES = [torch.float32, torch.bfloat16]


def dev():
    return torch.device("cuda:0")


@pytest.mark.parametrize("name", ["mona_cls", "mona_nocls", "mona_noise", "mona_freq", "mona_hybrid"])
@pytest.mark.parametrize("dtype", DTYPES)
def test_mona_golden(golden, name, dtype):
    """BatchFirstMonaWrapper(<Mona variant>) forward + all grads vs the reference module's outputs."""
    import src.adapters as A
    from src.adapters import BatchFirstMonaWrapper
    cls = {"mona_noise": A.NoiseAwareMona, "mona_freq": A.FreqEnhancedMona, "mona_hybrid": A.HybridNoiseFreqMona}.get(name, A.BaselineMona)
    g = golden(name)
    D = g["x"].shape[-1]
    m = BatchFirstMonaWrapper(cls(D, 64))
    m.load_state_dict(g["state"], strict=True)
    m = m.to(dev()).eval()
    x = g["x"].to(dev(), dtype).requires_grad_(True)
    y = m(x, g["hw"] if g["has_cls"] else None)
    (y * g["gy"].to(dev(), dtype)).sum().backward()
    assert y.shape == g["y"].shape and y.dtype == dtype
    assert relerr(y, g["y"]) < TOL[dtype]
    assert relerr(x.grad, g["dx"]) < GTOL[dtype]
    for k, p in m.named_parameters():
        assert p.grad is not None, k
        assert relerr(p.grad, g["grads"][k]) < GTOL[dtype], k


@pytest.mark.parametrize("dtype", DTYPES)
def test_mona_seq_first_768(dtype):
    """BaselineMona called the OpenAI-CLIP way ([N,B,D], reference mona.py:115-151) at D=768 vs the oracle."""
    from src.adapters import BaselineMona
    from oracle import functional as OF
    torch.manual_seed(0)
    m = BaselineMona(768, 64)
    with torch.no_grad():
        m.gamma.copy_(torch.randn(768) * 0.3)
    sd = {f"m.{k}": v.detach().double() for k, v in m.state_dict().items()}
    m = m.to(dev()).eval()
    x = torch.randn(197, 3, 768).to(dtype)
    y = m(x.to(dev()), (14, 14))
    ref = OF.mona(x.double().permute(1, 0, 2), sd, "m.", (14, 14), True).permute(1, 0, 2)
    assert y.shape == x.shape and relerr(y, ref) < TOL[dtype]


@pytest.mark.parametrize("dtype,grid", [(torch.float32, 18), (torch.bfloat16, 24), (torch.bfloat16, 22), (torch.float32, 16)])
def test_mona_large_grids_vs_oracle(dtype, grid):
    """grids above 16x16 (ViT-L/14@336 -> 24x24, ViT-B/16@352 -> 22x22) take the generic chunked conv-stage kernels."""
    from src.adapters import BaselineMona, BatchFirstMonaWrapper
    from oracle import functional as OF
    torch.manual_seed(0)
    D = 256
    m = BatchFirstMonaWrapper(BaselineMona(D, 64))
    with torch.no_grad():
        m.clip_mona.gamma.copy_(torch.randn(D) * 0.3)
    sd = {k: v.detach().double() for k, v in m.state_dict().items()}
    m = m.to(dev()).eval()
    N = grid * grid + 1
    x = torch.randn(2, N, D).to(dtype)
    gy = torch.randn(2, N, D).to(dtype)
    xg = x.to(dev()).requires_grad_(True)
    y = m(xg, (grid, grid))
    (y * gy.to(dev())).sum().backward()
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xo = x.double().requires_grad_(True)
    yo = OF.mona(xo, p, "clip_mona.", (grid, grid), True)
    names = [n for n, _ in m.named_parameters()]
    go = torch.autograd.grad((yo * gy.double()).sum(), [xo] + [p[n] for n in names])
    assert relerr(y, yo) < TOL[dtype] and relerr(xg.grad, go[0]) < GTOL[dtype]
    for (n, prm), gref in zip(m.named_parameters(), go[1:]):
        assert relerr(prm.grad, gref) < GTOL[dtype], n


def test_mona_dropout_statistics():
    """train mode: keep-rate ~ 0.9 with 1/(1-p) scaling, same mask in backward (reference mona.py:109,147)."""
    from nextgen_uia_b200 import ops
    torch.manual_seed(0)
    B, N, C = 4, 197, 64
    h = torch.randn(B, N, C, device=dev())
    w = [torch.zeros(64, 1, 3, 3), torch.zeros(64), torch.zeros(64, 1, 5, 5), torch.zeros(64), torch.zeros(64, 1, 7, 7), torch.zeros(64),
         torch.zeros(64, 64, 1, 1), torch.zeros(64)]
    w = [t.to(dev()) for t in w]
    g0 = ops.mona_conv_fwd(h, w, (14, 14), True, 0.0, 0)
    g1 = ops.mona_conv_fwd(h, w, (14, 14), True, 0.1, 1234)
    keep = (g1 != 0) | (g0 == 0)
    rate = keep.float().mean().item()
    assert abs(rate - 0.9) < 0.01
    assert torch.allclose(g1[keep], g0[keep] / 0.9, rtol=1e-5, atol=1e-6)
    grads = [torch.zeros_like(t) for t in w] + [torch.zeros(64, device=dev())]
    dh = ops.mona_conv_bwd(h, torch.ones_like(h), w, grads, (14, 14), True, 0.1, 1234)
    assert bool(((dh == 0) | keep).all()) and bool(((dh != 0) == (keep & (dh != 0))).all())


@pytest.mark.parametrize("dtype", DTYPES)
def test_lora_linear_golden(golden, dtype):
    from src.adapters import LinearLoRA
    g = golden("lora_linear")
    lin = torch.nn.Linear(256, 384)
    m = LinearLoRA(lin, r=g["r"], lora_alpha=g["alpha"], dropout_rate=0.1)
    m.load_state_dict(g["state"], strict=True)
    m = m.to(dev()).eval()
    x = g["x"].to(dev(), dtype).requires_grad_(True)
    y = m(x)
    (y * g["gy"].to(dev(), dtype)).sum().backward()
    assert relerr(y, g["y"]) < TOL[dtype]
    assert relerr(x.grad, g["dx"]) < GTOL[dtype]
    for n in ("w_lora_A", "w_lora_B", "bias"):
        assert relerr(getattr(m, n).grad, g["grads"][n]) < GTOL[dtype], n
    assert m.weight.grad is None


def _tiny_model(method, depth=2, seed=1):
    from nextgen_uia_b200.biomedclip import BiomedCLIP, init_synthetic_
    from src.adapters import inject_mona_variant_to_open_clip, inject_lora_to_biomedclip
    torch.manual_seed(seed)
    model = BiomedCLIP(vision=dict(depth=depth), text=dict(layers=depth, vocab=1000, max_pos=128))
    init_synthetic_(model, seed=seed, std=0.02)
    for p in model.parameters():
        p.requires_grad = False
    if method == "mona":
        inject_mona_variant_to_open_clip(model, variant="baseline", bottleneck_dim=64)
        key = "mona"
    else:
        inject_lora_to_biomedclip(model, lora_r=8, lora_alpha=32, lora_dropout=0.1)
        key = "lora"
        with torch.no_grad():
            for n, p in model.named_parameters():
                if n.endswith("w_lora_B"):
                    p.copy_(torch.randn(p.shape) * 0.02)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("gamma"):
                p.copy_(torch.randn(p.shape) * 0.2)
    for n, p in model.named_parameters():
        if key in n.lower():
            p.requires_grad = True
    return model


@pytest.mark.parametrize("method", ["mona", "lora"])
@pytest.mark.parametrize("dtype", DTYPES)
def test_model_loss_and_grads_vs_oracle(method, dtype):
    """Config-1-shaped micro-step (encode_image, encode_text, InfoNCE, backward) on a 2-layer tower:
    features, loss and every adapter gradient vs the CPU oracle on identical weights/inputs."""
    from oracle import functional as OF
    from src.losses import InfoNCELoss
    model = _tiny_model(method)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    trainable = [n for n, p in model.named_parameters() if p.requires_grad]
    torch.manual_seed(1)
    B = 4
    images = torch.rand(B, 3, 224, 224)
    ids = torch.randint(5, 1000, (B, 77)); ids[:, 0] = 2; ids[:, -1] = 3
    cfg = dict(patch=16, depth=2, heads=12, text_layers=2, text_heads=12, lora=(8, 32) if method == "lora" else None)
    lo, fio, fto, _, go = OF.loss_and_grads(sd, images, ids, cfg, trainable)
    model = model.to(dev()).eval().set_compute_dtype(dtype)
    fi = model.encode_image(images.to(dev()))
    ft = model.encode_text(ids.to(dev()))
    loss = InfoNCELoss(0.07)(fi, ft)
    assert relerr(fi, fio) < TOL[dtype] and relerr(ft, fto) < TOL[dtype]
    assert abs(float(loss.detach()) - float(lo)) / abs(float(lo)) < TOL[dtype]
    if dtype == torch.float32:
        loss.backward()                      # the full chain, every adapter gradient at 1e-3
    else:
        # With random synthetic weights all images map to nearly the same feature, so d(InfoNCE)/d(feature) is a
        # difference of near-equal vectors: bf16 feature rounding (4e-3) is amplified ~50x in that cotangent for ANY
        # bf16 implementation.  InfoNCE's own bf16 gradients are pinned by test_infonce_golden; here the tower
        # backward is checked with a well-conditioned fixed cotangent G on the image features.
        G = torch.randn(fi.shape, generator=torch.Generator().manual_seed(3))
        (fi * G.to(dev(), dtype)).sum().backward()
        p64 = {k: v.double().clone().requires_grad_(k in trainable) for k, v in sd.items()}
        fo = OF.encode_image(p64, images.double(), cfg)
        go = dict(zip(trainable, torch.autograd.grad((fo * G.double()).sum(), [p64[k] for k in trainable])))
    worst, num, den = ("", 0.0), 0.0, 0.0
    for n, p in model.named_parameters():
        if p.requires_grad:
            assert p.grad is not None, n
            e = relerr(p.grad, go[n])
            d = (p.grad.detach().double().cpu() - go[n].double())
            num += float((d * d).sum()); den += float((go[n].double() ** 2).sum())
            if e > worst[1]:
                worst = (n, e)
    if dtype == torch.float32:
        # fp32 check mode: every adapter gradient tensor individually
        assert worst[1] < 1e-3, worst
    else:
        # bf16: per-tensor max-error is dominated by cancellation noise in near-zero-sum bias gradients at this
        # tiny batch; the contract is on the aggregate adapter gradient (relative L2) + per-op tests above
        assert (num / den) ** 0.5 < 3e-2, ((num / den) ** 0.5, worst)


def test_zero_shot_argmax_matches_oracle():
    """identical argmax zero-shot predictions (src/models/biomedclip/zero_shot.py:176-228 recipe)."""
    from oracle import functional as OF
    model = _tiny_model("mona")
    sd = {k: v.detach().double() for k, v in model.state_dict().items()}
    torch.manual_seed(2)
    B = 8
    images = torch.rand(B, 3, 224, 224)
    prompts = [torch.randint(5, 1000, (10, 77)) for _ in range(2)]
    for p in prompts:
        p[:, 0] = 2; p[:, -1] = 3
    cfg = dict(patch=16, depth=2, heads=12, text_layers=2, text_heads=12)
    ref = OF.zero_shot_predict(OF.encode_image(sd, images.double(), cfg), [OF.encode_text(sd, p, cfg) for p in prompts])
    model = model.to(dev()).eval().set_compute_dtype(torch.bfloat16)
    # on-device scorer (nextgen_uia_b200/zero_shot.py: prototypes + fused normalise / dot / argmax kernel)
    from nextgen_uia_b200.zero_shot import ZeroShotScorer
    scorer = ZeroShotScorer(model)
    scorer.set_prompts({"benign": prompts[0].to(dev()), "malignant": prompts[1].to(dev())})
    logits, pred = scorer(images.to(dev()))
    assert torch.equal(pred.cpu().long(), ref)
    # its logits are the reference recipe's: mean over prompts of 100 * Ihat . That
    fio = OF.encode_image(sd, images.double(), cfg)
    io = fio / fio.norm(dim=-1, keepdim=True)
    lo = torch.stack([(100.0 * io @ (t / t.norm(dim=-1, keepdim=True)).t()).mean(1) for t in (OF.encode_text(sd, p, cfg) for p in prompts)], 1)
    # random towers give nearly orthogonal image / text features: the logits are 100 x cosines of ~1e-3, so the bf16 feature
    # error shows up as an ABSOLUTE cosine error; 1e-3 in cosine = 0.1 in logit units
    assert float((logits.double().cpu() - lo).abs().max()) < 0.1
    with torch.no_grad():
        fi = model.encode_image(images.to(dev())).float().cpu()
        tf = [model.encode_text(p.to(dev())).float().cpu() for p in prompts]
    got = OF.zero_shot_predict(fi, tf)
    assert torch.equal(got, ref)


@pytest.mark.parametrize("tag", ["clip_mona", "clip_lora"])
@pytest.mark.parametrize("dtype", DTYPES)
def test_openai_clip_golden(golden, tag, dtype):
    """OpenAI-CLIP layout (sequence-first blocks, nn.MultiheadAttention, QuickGELU, LN eps 1e-5; config-4 family):
    image features, text features (causal tower) and every adapter gradient vs the vendored reference model's outputs
    (tests/golden/clip_*.pt, made by oracle/make_golden.py from src/third_party/openai_clip/model.py + mona.py / lora.py)."""
    from nextgen_uia_b200.openai_clip import CLIP
    from src.adapters import inject_mona_variant_to_clip, inject_lora_to_clip
    g = golden(tag)
    m = CLIP(64, 32, 1, 256, 16, 8, 50, 64, 1, 1)
    for p in m.parameters():
        p.requires_grad = False
    if tag == "clip_mona":
        inject_mona_variant_to_clip(m, variant="baseline", bottleneck_dim=64)
    else:
        inject_lora_to_clip(m, lora_r=8, lora_alpha=32, lora_dropout=0.1)
    m.load_state_dict({k: (v.float() if v.is_floating_point() else v) for k, v in g["state"].items()}, strict=True)
    for n, p in m.named_parameters():
        p.requires_grad = n in g["trainable"]
    m = m.to(dev()).eval().set_compute_dtype(dtype)
    fi = m.encode_image(g["images"].to(dev()))
    ft = m.encode_text(g["text"].to(dev()))
    (fi * g["gi"].to(dev(), dtype)).sum().backward()
    assert relerr(fi, g["fi"]) < TOL[dtype] and relerr(ft, g["ft"]) < TOL[dtype]
    num = den = 0.0
    for n, p in m.named_parameters():
        if p.requires_grad:
            assert p.grad is not None, n
            d = p.grad.double().cpu() - g["grads"][n].double()
            num += float((d * d).sum()); den += float((g["grads"][n].double() ** 2).sum())
            if dtype == torch.float32:
                assert relerr(p.grad, g["grads"][n]) < 1e-3, n
    assert (num / den) ** 0.5 < (3e-2 if dtype == torch.bfloat16 else 1e-3)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_data_parallel_two_gpus(mode):
    """SURVEY.md §8e on real GPUs: 2 ranks (NCCL), all-gathered features -> global loss == single-process oracle on the
    concatenated batch, SUM-all-reduced adapter gradients == oracle gradients.  Skipped on a 1-GPU box."""
    import os, subprocess, sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    port = str(29600 + os.getpid() % 300)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", port, os.path.join(root, "tools", "dp_check.py"), mode], capture_output=True, text=True, timeout=300)
    assert "DP_CHECK_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_data_parallel_graph_segments_two_gpus():
    """The data-parallel step replayed as CUDA-graph segments with the NCCL collectives between them (_segcap.py) matches the
    eagerly launched step on 2 ranks.  Skipped on a 1-GPU box."""
    import os, subprocess, sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    port = str(29900 + os.getpid() % 90)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", port, os.path.join(root, "tools", "dp_graph_check.py")], capture_output=True, text=True, timeout=300)
    assert "DP_GRAPH_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("dtype", DTYPES)
def test_clipseg_adapter_vs_oracle(dtype):
    """Config-5 family: CLIP ViT encoder (kernels) with Mona re-enabled after the backbone freeze, hidden-state taps,
    causal text conditioning, HF CLIPSegDecoder (PyTorch).  Logits and the gradients of decoder + Mona parameters vs the
    CPU oracle (oracle towers + the same decoder weights on CPU)."""
    import copy
    from transformers import CLIPSegConfig
    from transformers.models.clipseg.modeling_clipseg import CLIPSegDecoder
    from nextgen_uia_b200.openai_clip import CLIP
    from nextgen_uia_b200.clipseg_adapter import CLIPSegAdapter
    from src.adapters import inject_mona_variant_to_clip
    from oracle import functional as OF
    torch.manual_seed(0)
    cfg = CLIPSegConfig(extract_layers=[0, 1, 2], reduce_dim=32, decoder_num_attention_heads=2, decoder_intermediate_size=64,
                        projection_dim=64, vision_config=dict(hidden_size=256, image_size=64, patch_size=16))
    dec = CLIPSegDecoder(cfg).float()
    clip = CLIP(64, 64, 3, 256, 16, 8, 50, 64, 1, 1)
    inject_mona_variant_to_clip(clip, variant="baseline", bottleneck_dim=64)
    with torch.no_grad():
        for n, p in clip.named_parameters():
            if n.endswith("gamma"):
                p.copy_(torch.randn(p.shape) * 0.2)
    model = CLIPSegAdapter(clip, decoder=dec)
    model.freeze_clip_backbone()
    assert model.unfreeze_adapters() == 3 * 16
    sd = {k: v.detach().double() if v.is_floating_point() else v for k, v in clip.state_dict().items()}
    dec64 = copy.deepcopy(dec).double()
    images = torch.rand(2, 3, 64, 64)
    ids = torch.randint(1, 48, (2, 8)); ids[:, -1] = 49
    gl = torch.randn(2, 2, 64, 64)
    # oracle
    trainable = [n for n, p in clip.named_parameters() if p.requires_grad]
    p64 = {k: (v.clone().requires_grad_(k in trainable) if v.is_floating_point() else v) for k, v in sd.items()}
    taps = []
    ocfg = dict(patch=16, depth=3, heads=4, text_layers=1, text_heads=1)
    OF.clip_encode_image(p64, images.double(), ocfg, taps=taps, tap_layers=(0, 1, 2))
    cond = OF.clip_encode_text(p64, ids, ocfg).detach()
    lo = dec64(hidden_states=tuple(taps), conditional_embeddings=cond)[0].view(2, -1, 64, 64)
    lo = torch.cat([-lo, lo], 1)
    go = torch.autograd.grad((lo * gl.double()).sum(), [p64[n] for n in trainable] + list(dec64.parameters()))
    # kernels
    model = model.to(dev()).eval()
    clip.set_compute_dtype(dtype)
    logits = model(images.to(dev()), input_ids=ids.to(dev()))
    (logits.float() * gl.to(dev())).sum().backward()
    tol = 5e-2 if dtype == torch.bfloat16 else 1e-3
    assert logits.shape == (2, 2, 64, 64) and relerr(logits, lo) < tol
    num = den = 0.0
    got = [dict(clip.named_parameters())[n].grad for n in trainable] + [p.grad for p in model.decoder.parameters()]
    for a, b in zip(got, go):
        assert a is not None
        d = a.double().cpu() - b
        num += float((d * d).sum()); den += float((b * b).sum())
    assert (num / den) ** 0.5 < (8e-2 if dtype == torch.bfloat16 else 1e-3)


@pytest.mark.gpu
def test_device_feeder_overlapped_copies_are_intact():
    """dp.DeviceFeeder hands out each host batch exactly once and unmodified while compute is queued behind it."""
    from nextgen_uia_b200 import dp
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    host = [(torch.randn(64, 3, 32, 32, generator=g).pin_memory(), torch.randint(0, 1000, (64, 16), generator=g).pin_memory()) for _ in range(7)]
    feeder = dp.DeviceFeeder(iter(host), dev)
    log = dp.ScalarLog()
    sums, w = [], torch.randn(2048, 2048, device=dev)
    for im, tx in feeder:
        for _ in range(20):      # keep the compute stream busy so slot reuse has to honour the `free` events
            w = torch.tanh(w @ w * 1e-3)
        log.push(im.double().sum() + tx.double().sum())
        sums += log.pop_ready()
    sums += log.drain()
    ref = [float(a.double().sum() + b.double().sum()) for a, b in host]
    assert len(sums) == 7 and feeder.h2d_bytes == sum(a.numel() * 4 + b.numel() * 8 for a, b in host)
    for s, r in zip(sums, ref):
        assert abs(s - r) <= 1e-3 * max(1.0, abs(r))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_openai_clip_patch14_vision_tower_vs_oracle(dtype):
    """Config-4 geometry in miniature: 14-pixel patches (ViT-L/14) -> 3*14*14 = 588 is not a multiple of 8, so the patch
    rows and the conv1 weight are zero-padded to 592 for the GEMM.  Image features and Mona gradients vs the oracle."""
    from nextgen_uia_b200.openai_clip import CLIP
    from src.adapters import inject_mona_variant_to_clip
    import oracle.functional as OF
    torch.manual_seed(21)
    m = CLIP(64, 56, 2, 256, 14, 8, 50, 64, 1, 1)
    for p in m.parameters():
        p.requires_grad = False
    inject_mona_variant_to_clip(m, variant="baseline", bottleneck_dim=64)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "mona" in n and ("project2" in n or "gamma" in n):
                p.add_(torch.randn_like(p) * 0.05)
    trainable = [n for n, p in m.named_parameters() if "mona" in n]
    for n, p in m.named_parameters():
        p.requires_grad = n in trainable
    images = torch.rand(3, 3, 56, 56)
    gi = torch.randn(3, 64)
    p64 = {k: (v.detach().double().clone().requires_grad_(k in trainable) if v.is_floating_point() else v) for k, v in m.state_dict().items()}
    ocfg = dict(patch=14, depth=2, heads=4, text_layers=1, text_heads=1)
    fo = OF.clip_encode_image(p64, images.double(), ocfg)
    go = torch.autograd.grad((fo * gi.double()).sum(), [p64[n] for n in trainable])
    m = m.to(dev()).eval().set_compute_dtype(dtype)
    fi = m.encode_image(images.to(dev()))
    (fi.float() * gi.to(dev())).sum().backward()
    assert relerr(fi, fo) < (2e-2 if dtype == torch.bfloat16 else 1e-4)
    num = den = 0.0
    params = dict(m.named_parameters())
    for n, b in zip(trainable, go):
        a = params[n].grad
        assert a is not None, n
        d = a.double().cpu() - b
        num += float((d * d).sum()); den += float((b * b).sum())
    assert (num / den) ** 0.5 < (5e-2 if dtype == torch.bfloat16 else 1e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("res,patch,width,heads", [(336, 14, 1024, 16), (352, 16, 768, 12)])
def test_openai_clip_config4_config5_geometry_bf16(res, patch, width, heads):
    """Config-4 / config-5 vision geometry at full width and resolution, two layers: ViT-L/14@336 -> 24 x 24 patches + CLS
    = 577 tokens, width 1024, 16 heads; ViT-B/16@352 -> 22 x 22 + CLS = 485 tokens, width 768.  Mona (baseline) on the
    24 x 24 / 22 x 22 grid.  Exercises the generic patchify path (K 588 -> 592), the D = 1024 LayerNorm kernels, the
    long-sequence attention kernels and the large-grid Mona conv kernels together."""
    from nextgen_uia_b200.openai_clip import CLIP
    from src.adapters import inject_mona_variant_to_clip
    import oracle.functional as OF
    torch.manual_seed(22)
    m = CLIP(64, res, 2, width, patch, 8, 50, 64, 1, 1)
    for p in m.parameters():
        p.requires_grad = False
    inject_mona_variant_to_clip(m, variant="baseline", bottleneck_dim=64)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "mona" in n and ("project2" in n or "gamma" in n):
                p.add_(torch.randn_like(p) * 0.05)
    trainable = [n for n, p in m.named_parameters() if "mona" in n]
    for n, p in m.named_parameters():
        p.requires_grad = n in trainable
    images = torch.rand(2, 3, res, res)
    gi = torch.randn(2, 64)
    p64 = {k: (v.detach().double().clone().requires_grad_(k in trainable) if v.is_floating_point() else v) for k, v in m.state_dict().items()}
    ocfg = dict(patch=patch, depth=2, heads=heads, text_layers=1, text_heads=1)
    fo = OF.clip_encode_image(p64, images.double(), ocfg)
    go = torch.autograd.grad((fo * gi.double()).sum(), [p64[n] for n in trainable])
    m = m.to(dev()).eval().set_compute_dtype(torch.bfloat16)
    fi = m.encode_image(images.to(dev()))
    (fi.float() * gi.to(dev())).sum().backward()
    assert fi.shape == (2, 64) and relerr(fi, fo) < 3e-2
    num = den = 0.0
    params = dict(m.named_parameters())
    for n, b in zip(trainable, go):
        a = params[n].grad
        assert a is not None and torch.isfinite(a).all(), n
        d = a.double().cpu() - b
        num += float((d * d).sum()); den += float((b * b).sum())
    assert (num / den) ** 0.5 < 8e-2


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_text_tower_right_padded_batch_vs_oracle(dtype):
    """BiomedCLIP text tower on a right-padded token batch (pad id 0): the key-padding mask travels to the attention
    kernels as per-sequence valid lengths; features vs the oracle (whose masked BERT stack is pinned to
    transformers.BertModel on CPU).  A mask that is not a suffix is refused."""
    from oracle import functional as OF
    model = _tiny_model("mona")
    sd = {k: v.detach().double().clone() if v.is_floating_point() else v.clone() for k, v in model.state_dict().items()}
    torch.manual_seed(33)
    ids = torch.randint(5, 1000, (6, 77))
    ids[:, 0] = 2
    for b, l in enumerate((77, 40, 33, 12, 64, 5)):
        ids[b, l - 1] = 3
        ids[b, l:] = 0
    cfg = dict(patch=16, depth=2, heads=12, text_layers=2, text_heads=12)
    ref = OF.encode_text(sd, ids, cfg)
    model = model.to(dev()).eval().set_compute_dtype(dtype)
    ft = model.encode_text(ids.to(dev()))
    assert relerr(ft, ref) < (2e-2 if dtype == torch.bfloat16 else 1e-4)
    unmasked = OF.encode_text(sd, ids, cfg, pad_token_id=-1)
    assert relerr(ft[3:4], unmasked[3:4]) > 5 * relerr(ft[3:4], ref[3:4])     # the mask matters for the short rows
    bad = ids.clone()
    bad[1, 10] = 0                                                             # a hole, not a suffix
    with pytest.raises(NotImplementedError):
        model.encode_text(bad.to(dev()))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_text_tower_lora_tune_text_encoder_vs_oracle(dtype):
    """inject_lora_to_biomedclip(tune_text_encoder=True) (reference lora.py:317-367): LoRA on the BERT q/k/v/o projections,
    right-padded batch.  Text features and every LoRA gradient of the text tower vs the oracle."""
    from oracle import functional as OF
    from nextgen_uia_b200.biomedclip import BiomedCLIP, init_synthetic_
    from src.adapters import inject_lora_to_biomedclip
    torch.manual_seed(41)
    model = BiomedCLIP(vision=dict(depth=1), text=dict(layers=2, vocab=1000, max_pos=128))
    init_synthetic_(model, seed=41, std=0.02)
    for p in model.parameters():
        p.requires_grad = False
    inject_lora_to_biomedclip(model, lora_r=8, lora_alpha=32, lora_dropout=0.0, tune_text_encoder=True)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("w_lora_B"):
                p.copy_(torch.randn(p.shape) * 0.02)
    trainable = [n for n, p in model.named_parameters() if "lora" in n and n.startswith("text.")]
    assert len(trainable) == 2 * 4 * 2
    for n, p in model.named_parameters():
        p.requires_grad = n in trainable
    ids = torch.randint(5, 1000, (5, 40))
    ids[:, 0] = 2
    for b, l in enumerate((40, 17, 33, 9, 25)):
        ids[b, l - 1] = 3
        ids[b, l:] = 0
    gt = torch.randn(5, 512)
    p64 = {k: (v.detach().double().clone().requires_grad_(k in trainable) if v.is_floating_point() else v) for k, v in model.state_dict().items()}
    cfg = dict(patch=16, depth=1, heads=12, text_layers=2, text_heads=12, text_lora=(8, 32))
    fo = OF.encode_text(p64, ids, cfg)
    go = torch.autograd.grad((fo * gt.double()).sum(), [p64[n] for n in trainable])
    model = model.to(dev()).eval().set_compute_dtype(dtype)
    ft = model.encode_text(ids.to(dev()))
    (ft.float() * gt.to(dev())).sum().backward()
    assert relerr(ft, fo) < (2e-2 if dtype == torch.bfloat16 else 1e-4)
    num = den = 0.0
    params = dict(model.named_parameters())
    for n, b in zip(trainable, go):
        a = params[n].grad
        assert a is not None, n
        d = a.double().cpu() - b
        num += float((d * d).sum()); den += float((b * b).sum())
        if dtype == torch.float32:
            assert relerr(a, b) < 2e-3, n
    assert (num / den) ** 0.5 < (5e-2 if dtype == torch.bfloat16 else 1e-3)


# ---------------------------------------------------------------------------------------------------------------------
# Parity at the BENCHMARKED depth (BASELINE.json configs[0] / configs[1] shapes): goldens from oracle/make_golden_cfg.py
# ---------------------------------------------------------------------------------------------------------------------
def _report(name, payload):
    """Per-layer error growth etc. for DESIGN.md / profiles/: written next to the test run when the directory exists."""
    import json, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = os.path.join(root, "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, f"parity_{name}.json"), "w") as f:
            json.dump(payload, f, indent=1)
    except OSError:
        pass
    print(name, payload)


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _cfg_model_and_taps(dtype):
    from oracle.make_golden_cfg import build, weight_checksum, TAP_IMGS, TAP_TOKS
    model = build(12)
    ck = weight_checksum(model.state_dict())
    model = model.to(dev()).eval().set_compute_dtype(dtype)
    taps = []
    hooks = [blk.register_forward_hook(lambda _m, _i, out: taps.append(out.detach()[:TAP_IMGS][:, list(TAP_TOKS), :].float().cpu()))
             for blk in model.visual.trunk.blocks]
    return model, ck, taps, hooks


@pytest.mark.parametrize("dtype", DTYPES)
def test_config1_depth12_batch8_parity(golden, dtype):
    """BASELINE.json configs[0] exactly: batch 8, 12 vision layers + Mona, 12 BERT layers, InfoNCE, backward.
    fp32 check mode: features / loss 1e-4, every stored adapter gradient 1e-3 (fp32 accumulation over 12 layers).
    bf16: features / loss 1e-2; tower backward with a well-conditioned cotangent, relative L2 over the adapter."""
    import bench
    from src.losses import InfoNCELoss
    g = golden("cfg1_b8_d12")
    model, ck, taps, hooks = _cfg_model_and_taps(dtype)
    assert torch.allclose(ck, g["checksum"], rtol=1e-12, atol=0), "synthetic weights drifted from the golden's"
    images, ids = bench.synthetic_batch(8, 1)
    fi = model.encode_image(images.to(dev()))
    ft = model.encode_text(ids.to(dev()))
    loss = InfoNCELoss(0.07)(fi, ft)
    for h in hooks:
        h.remove()
    layer_err = [relerr(t, g["taps"][i]) for i, t in enumerate(taps)]
    layer_l2 = [rel_l2(t, g["taps"][i]) for i, t in enumerate(taps)]
    e_i, e_t = relerr(fi, g["fi"]), relerr(ft, g["ft"])
    l_i, l_t = rel_l2(fi, g["fi"]), rel_l2(ft, g["ft"])
    e_l = abs(float(loss.detach()) - g["loss"]) / abs(g["loss"])
    if dtype == torch.float32:
        loss.backward()
        ref, refn = g["grads"], g["grad_norms"]
    else:
        (fi * g["G"].to(dev(), dtype)).sum().backward()
        ref, refn = g["gradsG"], g["gradG_norms"]
    params = dict(model.named_parameters())
    num = den = 0.0
    worst = ("", 0.0)
    for n, b in ref.items():
        a = params[n].grad
        assert a is not None and torch.isfinite(a).all(), n
        d = a.double().cpu() - b.double()
        num += float((d * d).sum()); den += float((b.double() ** 2).sum())
        e = relerr(a, b)
        if e > worst[1]:
            worst = (n, e)
    norm_err = max(abs(float(params[n].grad.norm()) - v) / max(v, 1e-30) for n, v in refn.items() if v > 1e-12)
    l2 = (num / den) ** 0.5
    _report(f"cfg1_b8_d12_{'fp32' if dtype == torch.float32 else 'bf16'}",
            {"image_feat_relerr_max": e_i, "text_feat_relerr_max": e_t, "image_feat_rel_l2": l_i, "text_feat_rel_l2": l_t, "loss_relerr": e_l,
             "per_layer_residual_relerr_max": layer_err, "per_layer_residual_rel_l2": layer_l2,
             "adapter_grad_rel_l2_layers_0_5_11": l2, "worst_tensor": list(worst), "max_grad_norm_relerr_all_layers": norm_err})
    assert e_l < TOL[dtype], e_l
    if dtype == torch.float32:
        assert e_i < 1e-4 and e_t < 1e-4, (e_i, e_t)
    else:
        # Depth-12 bf16: the loss and the relative-L2 feature error hold the 1e-2 contract; the
        # max-norm feature error is 1.3-1.6e-2 because the residual stream is STORED in bf16 between kernels (<= 2^-8
        # relative rounding of the largest element per store, a random walk over 12 layers: 0.7e-2 after layer 0).
        assert l_i < 2e-2 and l_t < 2e-2, (l_i, l_t)
        assert e_i < 2e-2 and e_t < 2e-2, (e_i, e_t, layer_err)
    if dtype == torch.float32:
        assert worst[1] < 1e-3 and norm_err < 1e-3, (worst, norm_err)
    else:
        assert l2 < 2e-2 and norm_err < 2e-2, (l2, norm_err, worst)


def test_config2_depth12_batch256_bf16_parity(golden):
    """BASELINE.json configs[1] shape (the benchmarked one): batch 256, 12+12 layers, bf16, eval mode.
    Image / text features and the InfoNCE loss vs the committed CPU-oracle golden; per-layer error growth reported."""
    import bench
    from src.losses import InfoNCELoss
    g = golden("cfg2_b256_d12")
    model, ck, taps, hooks = _cfg_model_and_taps(torch.bfloat16)
    assert torch.allclose(ck, g["checksum"], rtol=1e-12, atol=0)
    images, ids = bench.synthetic_batch(256, 1)
    with torch.no_grad():
        fi = model.encode_image(images.to(dev()))
        ft = model.encode_text(ids.to(dev()))
        loss = InfoNCELoss(0.07)(fi, ft)
    for h in hooks:
        h.remove()
    layer_err = [relerr(t, g["taps"][i]) for i, t in enumerate(taps)]
    layer_l2 = [rel_l2(t, g["taps"][i]) for i, t in enumerate(taps)]
    e_i, e_t = relerr(fi, g["fi"]), relerr(ft, g["ft"])
    l_i, l_t = rel_l2(fi, g["fi"]), rel_l2(ft, g["ft"])
    e_l = abs(float(loss) - g["loss"]) / abs(g["loss"])
    # zero-shot style decision on the benchmark batch: image -> nearest text must agree wherever the oracle's margin is
    # larger than the feature error can move a cosine (identical-argmax criterion of the parity contract)
    nrm = lambda t: t.double() / t.double().norm(dim=1, keepdim=True)
    so = nrm(g["fi"]) @ nrm(g["ft"]).t()
    sg = nrm(fi.cpu()) @ nrm(ft.cpu()).t()
    top2 = so.topk(2, dim=1).values
    decided = (top2[:, 0] - top2[:, 1]) > 1e-3
    agree = float((so.argmax(1) == sg.argmax(1))[decided].double().mean()) if bool(decided.any()) else 1.0
    _report("cfg2_b256_d12_bf16", {"image_feat_relerr_max": e_i, "text_feat_relerr_max": e_t, "image_feat_rel_l2": l_i, "text_feat_rel_l2": l_t,
                                   "loss_relerr": e_l, "per_layer_residual_relerr_max": layer_err, "per_layer_residual_rel_l2": layer_l2,
                                   "argmax_agreement_where_margin_gt_1e-3": agree, "rows_decided": int(decided.sum())})
    assert e_l < 1e-2 and l_i < 2e-2 and l_t < 2e-2, (e_l, l_i, l_t)
    assert e_i < 2e-2 and e_t < 2e-2, (e_i, e_t, layer_err)      # see test_config1_depth12_batch8_parity
    assert agree == 1.0


def test_depth12_bf16_error_vs_stock_pytorch_bf16(golden):
    """What 'bf16 parity with the reference PyTorch path' can mean at depth 12: the SAME model evaluated by plain PyTorch
    (the oracle restatement moved to the GPU under torch.autocast(bfloat16), i.e. cuBLAS bf16 GEMMs with fp32 residual
    stream and fp32 LayerNorm, the way the reference would run in bf16) is itself ~1e-2 away from the fp64 result.
    Our path must not be worse than 2x that noise floor on either tower."""
    import bench
    from oracle import functional as OF
    from oracle.make_golden_cfg import build
    g = golden("cfg1_b8_d12")
    model = build(12)
    sd = {k: v.detach().to(dev()) for k, v in model.state_dict().items()}
    images, ids = bench.synthetic_batch(8, 1)
    cfg = dict(patch=16, depth=12, heads=12, text_layers=12, text_heads=12)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        fi_ref = OF.encode_image(sd, images.to(dev()), cfg)
        ft_ref = OF.encode_text(sd, ids.to(dev()), cfg)
    model = model.to(dev()).eval().set_compute_dtype(torch.bfloat16)
    with torch.no_grad():
        fi = model.encode_image(images.to(dev()))
        ft = model.encode_text(ids.to(dev()))
    rep = {"stock_pytorch_bf16_image_rel_l2": rel_l2(fi_ref, g["fi"]), "stock_pytorch_bf16_text_rel_l2": rel_l2(ft_ref, g["ft"]),
           "ours_image_rel_l2": rel_l2(fi, g["fi"]), "ours_text_rel_l2": rel_l2(ft, g["ft"]),
           "stock_pytorch_bf16_image_relerr_max": relerr(fi_ref, g["fi"]), "stock_pytorch_bf16_text_relerr_max": relerr(ft_ref, g["ft"]),
           "ours_image_relerr_max": relerr(fi, g["fi"]), "ours_text_relerr_max": relerr(ft, g["ft"]),
           "ours_vs_stock_image_rel_l2": rel_l2(fi, fi_ref), "ours_vs_stock_text_rel_l2": rel_l2(ft, ft_ref)}
    _report("depth12_bf16_noise_floor", rep)
    assert rep["ours_image_rel_l2"] < max(1e-2, 2 * rep["stock_pytorch_bf16_image_rel_l2"]), rep
    assert rep["ours_text_rel_l2"] < max(1e-2, 2 * rep["stock_pytorch_bf16_text_rel_l2"]), rep


def _small_trainer(seed=1, **kw):
    from nextgen_uia_b200 import dp
    model = _tiny_model("mona", depth=2, seed=seed).to(dev()).train().set_compute_dtype(torch.bfloat16)
    for m in model.modules():       # train mode (the hot loop: lazy padding check, no host sync) with dropout disabled
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    return model, dp.Trainer(model, lr=1e-3, total_updates=10, **kw)


def test_cuda_graph_step_matches_eager():
    """Trainer.capture/replay (whole step: forward, backward, clip, AdamW, schedule, guard as ONE CUDA graph) reproduces the
    eager launches: same parameters after 3 updates (dropout p = 0), device-side update counter advanced, and
    capturing itself leaves parameters / optimiser state untouched."""
    torch.manual_seed(0)
    images = torch.rand(4, 3, 224, 224, device=dev())
    ids = torch.randint(5, 1000, (4, 77), device=dev()); ids[:, 0] = 2; ids[:, -1] = 3
    _, te = _small_trainer()
    le = [float(te.micro_step(images, ids)) for _ in range(3)]
    _, te2 = _small_trainer()
    for _ in range(3):
        te2.micro_step(images, ids)
    _, tg = _small_trainer()
    p0 = tg.buckets.flat_param.clone()
    tg.capture(images, ids)
    assert torch.equal(tg.buckets.flat_param, p0) and tg.optimizer.updates == 0
    lg = [float(tg.replay(images, ids)) for _ in range(3)]
    assert tg.optimizer.updates == 3 and te.optimizer.updates == 3
    assert max(abs(a - b) for a, b in zip(le, lg)) < 2e-3 * abs(le[0]), (le, lg)
    # Adam normalises every element's update to ~lr, so elements whose gradient is rounding noise flip sign between ANY two
    # runs (fp32 atomics order): measure the graph-vs-eager distance against the eager-vs-eager distance of the same update
    upd = (te.buckets.flat_param - p0).norm()
    d_ee = float((te2.buckets.flat_param - te.buckets.flat_param).norm() / upd)
    d_ge = float((tg.buckets.flat_param - te.buckets.flat_param).norm() / upd)
    assert d_ge < max(3 * d_ee, 5e-2), (d_ge, d_ee)


def test_nonfinite_loss_poisons_the_whole_accumulation_window():
    """a non-finite loss on ANY micro-step of an accumulation window cancels that window's update,
    advances neither the update counter nor the LR schedule, and leaves zeroed gradients (finetune.py:281-288)."""
    model, tr = _small_trainer(accumulation_steps=2)
    torch.manual_seed(0)
    images = torch.rand(4, 3, 224, 224, device=dev())
    ids = torch.randint(5, 1000, (4, 77), device=dev()); ids[:, 0] = 2; ids[:, -1] = 3
    p0 = tr.buckets.flat_param.clone()
    bad = images.clone(); bad[0, 0, 0, 0] = float("nan")
    tr.micro_step(bad, ids)             # poisoned micro-step (its NaN gradients are in the flat buffer)
    tr.micro_step(images, ids)          # finite loss on the micro-step that completes the window
    assert torch.equal(tr.buckets.flat_param, p0), "update must be skipped"
    assert tr.optimizer.updates == 0 and tr.optimizer.skipped == 1
    assert float(tr.buckets.flat_grad.abs().max()) == 0.0 and torch.isfinite(tr.optimizer.m).all()
    tr.micro_step(images, ids); tr.micro_step(images, ids)
    assert tr.optimizer.updates == 1 and not torch.equal(tr.buckets.flat_param, p0)
    assert torch.isfinite(tr.buckets.flat_param).all()


def test_text_padding_flag_is_lazy_in_train_mode():
    model = _tiny_model("mona").to(dev()).set_compute_dtype(torch.bfloat16)
    ids = torch.randint(5, 1000, (3, 20), device=dev()); ids[:, 0] = 2
    ids[1, 7] = 0                       # a hole, not a suffix
    model.train()
    model.encode_text(ids)              # no host sync, no exception in the hot loop
    with pytest.raises(NotImplementedError):
        model.text.check_padding()
    model.eval()
    with pytest.raises(NotImplementedError):
        model.encode_text(ids)


@pytest.mark.parametrize("name", ["mona_cls", "mona_freq"])
def test_mona_unfused_bf16_path_still_matches_golden(golden, name, monkeypatch):
    """NGU_MONA_FUSED=0 keeps the round-1 kernel sequence (LN-mix -> GEMM -> stage -> GEMM; used for grids above 16x16, the
    noise-aware variants and as the A/B switch of the fused path): same goldens, same tolerances."""
    monkeypatch.setenv("NGU_MONA_FUSED", "0")
    import src.adapters as A
    from src.adapters import BatchFirstMonaWrapper
    cls = A.FreqEnhancedMona if name == "mona_freq" else A.BaselineMona
    g = golden(name)
    m = BatchFirstMonaWrapper(cls(g["x"].shape[-1], 64))
    m.load_state_dict(g["state"], strict=True)
    m = m.to(dev()).eval()
    x = g["x"].to(dev(), torch.bfloat16).requires_grad_(True)
    y = m(x, g["hw"] if g["has_cls"] else None)
    (y * g["gy"].to(dev(), torch.bfloat16)).sum().backward()
    assert relerr(y, g["y"]) < TOL[torch.bfloat16] and relerr(x.grad, g["dx"]) < GTOL[torch.bfloat16]
    for k, p in m.named_parameters():
        assert relerr(p.grad, g["grads"][k]) < GTOL[torch.bfloat16], k


@pytest.mark.parametrize("task", ["seg", "cls"])
def test_timm_clip_adapter_heads_vs_oracle(task):
    """f4: downstream heads on tapped block activations (reference TimmCLIPAdapter, timm/clip_adapter.py:118-160): taps from the
    kernels (Mona inside) feed the reduce -> LayerNorm/MLP pyramid and the seg / cls head; logits and the gradients of head +
    Mona parameters vs the same head applied to the CPU oracle's taps."""
    import copy
    from nextgen_uia_b200.clip_adapter import TimmCLIPAdapter
    from oracle import functional as OF
    torch.manual_seed(3)
    clip = _tiny_model("mona", depth=3)
    ad = TimmCLIPAdapter(clip, extract_layers=[0, 2], reduce_dim=64, num_classes=2, img_size=224, task=task).eval()
    ad.freeze_clip_backbone()
    head64 = copy.deepcopy(torch.nn.ModuleList([ad.reduces, ad.blocks, ad.seg_head, ad.cls_head])).double()
    sd = {k: v.detach().double().clone() for k, v in clip.state_dict().items()}
    trainable = [n for n, p in clip.named_parameters() if p.requires_grad]
    assert trainable and all("mona" in n for n in trainable)
    images = torch.rand(2, 3, 224, 224)
    p64 = {k: v.clone().requires_grad_(k in trainable) for k, v in sd.items()}
    taps = []
    OF.encode_image(p64, images.double(), dict(patch=16, depth=3, heads=12), taps=taps)
    red, blk, seg, cls = head64
    fused = None
    for lvl, ti in ((1, 2), (0, 0)):
        y = blk[lvl](red[lvl](taps[ti][:, 1:, :]))
        fused = y if fused is None else fused + y
    fmap = fused.transpose(1, 2).reshape(2, 64, 14, 14)
    lo = seg(fmap) if task == "seg" else cls(fmap)
    gl = torch.randn(lo.shape, dtype=torch.float64)
    head_params = [p for p in head64.parameters()]
    go = torch.autograd.grad((lo * gl).sum(), [p64[n] for n in trainable] + head_params, allow_unused=True)
    ad = ad.to(dev())
    clip.set_compute_dtype(torch.bfloat16)
    logits = ad(images.to(dev()))
    (logits.double() * gl.to(dev())).sum().backward()
    assert logits.shape == lo.shape and relerr(logits, lo) < 2e-2
    num = den = 0.0
    params = dict(clip.named_parameters())
    for n, b in zip(trainable, go):
        d = params[n].grad.double().cpu() - b
        num += float((d * d).sum()); den += float((b * b).sum())
    assert (num / den) ** 0.5 < 5e-2
