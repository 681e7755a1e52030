"""CPU: host-side logic of the drop-in boundary (no kernels run): injection, freeze/thaw by name,
state-dict key format, error behaviour, and the data-parallel gradient plumbing over gloo."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model(depth=2):
    from nextgen_uia_b200.biomedclip import BiomedCLIP
    return BiomedCLIP(vision=dict(depth=depth), text=dict(layers=1, vocab=100, max_pos=80))


def test_mona_injection_keys_and_thaw():
    from src.adapters import inject_mona_variant_to_open_clip
    m = _model()
    for p in m.parameters():
        p.requires_grad = False
    m, n = inject_mona_variant_to_open_clip(m, variant="baseline", bottleneck_dim=64)
    assert n == 2
    keys = [k for k in m.state_dict() if "mona" in k]
    # checkpoint wire format
    want = ["gamma", "gammax", "project1.weight", "project1.bias", "project2.weight", "project2.bias",
            "adapter_conv.conv1.weight", "adapter_conv.conv1.bias", "adapter_conv.conv2.weight", "adapter_conv.conv2.bias",
            "adapter_conv.conv3.weight", "adapter_conv.conv3.bias", "adapter_conv.projector.weight", "adapter_conv.projector.bias",
            "norm.weight", "norm.bias"]
    assert keys[:16] == [f"visual.trunk.blocks.0.mona.clip_mona.{w}" for w in want]
    sd = m.state_dict()
    assert tuple(sd["visual.trunk.blocks.0.mona.clip_mona.adapter_conv.conv3.weight"].shape) == (64, 1, 7, 7)
    assert tuple(sd["visual.trunk.blocks.0.mona.clip_mona.adapter_conv.projector.weight"].shape) == (64, 64, 1, 1)
    for name, p in m.named_parameters():
        if "mona" in name.lower():
            p.requires_grad = True
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == 2 * 111872  # 111,872 / layer


def test_unknown_variant_raises_like_reference():
    from src.adapters import inject_mona_variant_to_open_clip, inject_mona_variant_to_clip
    with pytest.raises(ValueError, match="Unknown variant"):
        inject_mona_variant_to_open_clip(_model(), variant="fractional")
    with pytest.raises(ValueError, match="Unknown variant"):
        inject_mona_variant_to_clip(_model(), variant="nope")


def test_injection_is_silent_noop_without_expected_tree(capsys):
    from src.adapters import inject_mona_variant_to_open_clip, inject_lora_to_biomedclip, inject_lora_to_clip
    dummy = torch.nn.Linear(2, 2)
    _, n = inject_mona_variant_to_open_clip(dummy, variant="baseline")
    assert n == 0
    assert inject_lora_to_biomedclip(dummy)[1] == 0 and inject_lora_to_clip(dummy)[1] == 0
    assert "✓" in capsys.readouterr().out


def test_lora_injection_and_trainable_bias_quirk():
    from src.adapters import inject_lora_to_biomedclip, LinearLoRA
    m = _model()
    for p in m.parameters():
        p.requires_grad = False
    m, n = inject_lora_to_biomedclip(m, lora_r=8, lora_alpha=32, lora_dropout=0.1)
    assert n == 2
    blk = m.visual.trunk.blocks[0]
    assert isinstance(blk.attn.qkv, LinearLoRA) and isinstance(blk.attn.proj, LinearLoRA)
    assert tuple(blk.attn.qkv.w_lora_A.shape) == (8, 768) and tuple(blk.attn.qkv.w_lora_B.shape) == (2304, 8)
    # reference quirk: the fresh nn.Linear's bias stays trainable, the weight is frozen
    assert blk.attn.qkv.bias.requires_grad and not blk.attn.qkv.weight.requires_grad
    trainable = sum(p.numel() for p in m.parameters() if p.requires_grad)
    assert trainable == 2 * (8 * 768 + 2304 * 8 + 2304 + 8 * 768 + 768 * 8 + 768)


def test_patched_forward_passes_kwargs():
    from src.adapters import inject_mona_variant_to_open_clip
    m = _model(1)
    blk = m.visual.trunk.blocks[0]
    seen = {}
    blk.forward = lambda x, **kw: seen.update(kw) or x
    inject_mona_variant_to_open_clip(m, variant="baseline")
    class Plus1(torch.nn.Module):
        def forward(self, x, hw):
            return x + 1
    blk.mona = Plus1()
    out = blk.forward(torch.zeros(1), attn_mask="k")
    assert seen == {"attn_mask": "k"} and float(out) == 1.0


def test_full_finetune_is_rejected_not_silently_wrong():
    from nextgen_uia_b200.linear import frozen_copies
    w = torch.nn.Parameter(torch.zeros(4, 4))
    with pytest.raises(NotImplementedError):
        frozen_copies(w, torch.bfloat16)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` prints one JSON line with the contract keys (tiny depth so it is quick)."""
    import json
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                                   "--depth", "1", "--cpu-batch", "2"], text=True, cwd=ROOT)
    line = json.loads(out.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "cpu_baseline", "e2e", "config", "higher_is_better"):
        assert k in line
    # "reference" when oracle/_ref holds the reference's own modules (oracle/build_ref.py, built where /root/reference exists and
    # shipped with the snapshot), "port" otherwise
    want = "reference" if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "mona.py")) else "port"
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == want and line["value"] > 0


GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
from nextgen_uia_b200.dp import GradBuckets
torch.manual_seed(0)
ps = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(5))]
gb = GradBuckets(ps, lambda p: 0)
rank = dist.get_rank()
# per-rank gradient contributions accumulate into the flat bucket views; SUM all-reduce makes them global
# backward through autograd: the post-accumulate hooks launch the bucket all-reduce when the bucket is complete
loss = (ps[0] * float(rank + 1)).sum() + (ps[1] * float(10 * (rank + 1))).sum()
loss.backward()
gb.wait()
assert torch.allclose(ps[0].grad, torch.full((3, 4), 3.0)) and torch.allclose(ps[1].grad, torch.full((5,), 30.0))
# label offsets of the global logit matrix: rank r owns rows [r*Bl, (r+1)*Bl)
Bl = 4
allf = [torch.zeros(Bl, 2) for _ in range(2)]
dist.all_gather(allf, torch.full((Bl, 2), float(rank)))
g = torch.cat(allf)
assert g[rank * Bl:(rank + 1) * Bl].eq(rank).all()
gb.zero()
assert float(ps[0].grad.abs().sum()) == 0.0
dist.destroy_process_group()
print("ok")
"""


def test_dp_gradient_plumbing_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(GLOO_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    for p in procs:
        out, err = p.communicate(timeout=120)
        assert p.returncode == 0 and "ok" in out, err[-2000:]


def test_device_feeder_and_scalar_log_cpu():
    """Host-side input pipeline and loss log (dp.DeviceFeeder / dp.ScalarLog): order, byte accounting, exhaustion."""
    from nextgen_uia_b200 import dp
    batches = [(torch.full((4, 3), float(i)), torch.full((4,), i, dtype=torch.int64)) for i in range(5)]
    feeder = dp.DeviceFeeder(iter(batches), "cpu")
    seen = [(int(a[0, 0]), int(b[0])) for a, b in feeder]
    assert seen == [(i, i) for i in range(5)]
    assert feeder.h2d_bytes == 5 * (4 * 3 * 4 + 4 * 8)
    log = dp.ScalarLog()
    got = []
    for i in range(5):
        log.push(torch.tensor(float(i)))
        got += log.pop_ready()
    assert got == [0.0, 1.0, 2.0, 3.0]
    assert log.drain() == [4.0]


def test_src_nets_shim_importable_with_reference_signatures():
    """src.nets keeps conv_layer, linear_layer, CoordConv, Projector, FPN_AD, CARAFE, Masker, gumbel_softmax
    importable with the reference's signatures (dead code there too); plus the src.third_party re-export paths."""
    import inspect
    from src.nets.layers import conv_layer, linear_layer, CoordConv, Projector, FPN_AD
    from src.nets.carafe import CARAFE
    from src.nets.masker import Masker, gumbel_softmax
    assert list(inspect.signature(conv_layer).parameters) == ["in_dim", "out_dim", "kernel_size", "padding", "stride"]
    assert list(inspect.signature(linear_layer).parameters) == ["in_dim", "out_dim", "bias"]
    assert list(inspect.signature(CARAFE.__init__).parameters)[1:] == ["inC", "outC", "kernel_size", "up_factor"]
    assert list(inspect.signature(Projector.__init__).parameters)[1:] == ["word_dim", "in_dim", "kernel_size"]
    assert list(inspect.signature(Masker.__init__).parameters)[1:] == ["in_dim", "outdim"]
    assert list(inspect.signature(FPN_AD.__init__).parameters)[1:] == ["in_channels", "out_channels"]
    x = torch.randn(2, 16, 5, 7)
    assert CARAFE(16, 8)(x).shape == (2, 8, 10, 14)
    assert CoordConv(16, 4)(x).shape == (2, 4, 5, 7)
    assert Masker(16, 16).eval()(x).shape == (2, 16, 5, 7)
    assert Projector(32, 8).eval()(torch.randn(2, 16, 3, 3), torch.randn(2, 32)).shape == (2, 1, 48, 48)
    p = gumbel_softmax(torch.randn(4, 5))
    assert torch.allclose(p.sum(-1), torch.ones(4))
    assert "masker5.conv.weight" in FPN_AD().state_dict() and "carafe.encoder.weight" in FPN_AD().state_dict()
    from src.third_party.timm.clip_adapter import TimmCLIPAdapter  # noqa: F401
    from src.third_party.openai_clip.clip_adapter import CLIPAdapter  # noqa: F401
    from src.third_party.openai_clip.clipseg_adapter import CLIPSegAdapter  # noqa: F401
    from src.third_party.openai_clip.model import CLIP  # noqa: F401


def test_only_the_causal_attention_mask_is_accepted():
    """PlainMultiheadAttentionLoRA / ResidualAttentionBlock run exactly one mask, CLIP's causal text mask (reference
    model.py:344-350); anything else must be refused, not mis-applied (reference lora.py:178-190 forwards the mask to SDPA)."""
    import pytest, torch
    from nextgen_uia_b200.linear import require_causal_mask
    L = 7
    m = torch.full((L, L), float("-inf")).triu_(1)
    require_causal_mask(m, L)                      # the reference's build_attention_mask
    require_causal_mask(m.to(torch.bfloat16), L)
    for bad in (torch.zeros(L, L), torch.ones(L, L, dtype=torch.bool).tril(), m[:, :5], m.clone().fill_(0).fill_diagonal_(float("-inf")),
                torch.full((L, L), -1e4).triu_(1)):
        with pytest.raises(NotImplementedError):
            require_causal_mask(bad, L)
    with pytest.raises(NotImplementedError):
        require_causal_mask(m, L + 1)


def test_segmented_capture_bookkeeping_without_cuda():
    """_segcap: outside a capture collective(fn) just runs fn; a replay walks graph segments and collectives in recording order."""
    from nextgen_uia_b200 import _segcap
    calls = []
    _segcap.collective(lambda: calls.append("eager"))
    assert calls == ["eager"] and _segcap.ACTIVE is None

    class FakeGraph:                      # stands in for torch.cuda.CUDAGraph (no GPU here)
        def __init__(self, name):
            self.name = name

        def replay(self):
            calls.append(self.name)

    seg = _segcap.SegmentedCapture()
    seg.items = ["g0", (lambda: calls.append("all_gather")), "g1", (lambda: calls.append("all_reduce")), "g2"]
    import torch
    real = torch.cuda.CUDAGraph
    try:
        torch.cuda.CUDAGraph = FakeGraph
        seg.items = [FakeGraph(x) if isinstance(x, str) else x for x in seg.items]
        del calls[:]
        seg.replay()
        assert calls == ["g0", "all_gather", "g1", "all_reduce", "g2"] and seg.segments == 3
    finally:
        torch.cuda.CUDAGraph = real
