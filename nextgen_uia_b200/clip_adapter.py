"""Downstream heads on tapped block activations — mirror of the reference's TimmCLIPAdapter
(src/third_party/timm/clip_adapter.py:6-200) and CLIPAdapter (src/third_party/openai_clip/clip_adapter.py:6-170).

Same class names, constructor arguments, attribute names (`clip_model`, `reduces`, `blocks`, `seg_head`, `cls_head`: so head
checkpoints interchange), methods (`extract_vit_features`, `forward`, `freeze_clip_backbone`) and output shapes.  The encoder
blocks (+ Mona / LoRA) run on the B200 kernels; the small trainable head (per-level Linear -> LayerNorm/MLP, pyramid sum,
bilinear up-sampling + 1x1 conv, or pooled classifier) is ordinary PyTorch, like the reference's.
"""
import math

import torch
import torch.nn as nn


def _level_mlp(dim):
    return nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, dim), nn.GELU(), nn.Linear(dim, dim))


class _PyramidHeadBase(nn.Module):
    """Shared machinery: taps after the blocks listed in `extract_layers`, deep-to-shallow pyramid sum, task head."""

    def __init__(self, clip_model, feature_dim, extract_layers, reduce_dim, num_classes, img_size, task, cls_head):
        super().__init__()
        if task not in ("seg", "cls"):
            raise ValueError(f"Invalid task type: {task}")
        self.clip_model = clip_model
        self.extract_layers = list(extract_layers)
        self.reduce_dim, self.num_classes, self.task, self.feature_dim = reduce_dim, num_classes, task, feature_dim
        n = len(self.extract_layers)
        self.reduces = nn.ModuleList(nn.Linear(feature_dim, reduce_dim) for _ in range(n))
        self.blocks = nn.ModuleList(_level_mlp(reduce_dim) for _ in range(n))
        self.seg_head = nn.Sequential(nn.Upsample((img_size, img_size), mode="bilinear", align_corners=False),
                                      nn.Conv2d(reduce_dim, num_classes, kernel_size=1))
        self.cls_head = cls_head

    # ---- taps: batch-first [B, 1 + grid^2, D] hidden states after the selected blocks -------------------------------
    def _tower_taps(self, x):
        vis = self.clip_model.visual
        taps = []
        if hasattr(vis, "trunk"):                               # timm / open_clip layout: blocks take [B, N, D]
            h = vis.trunk.embed(x)
            for i, blk in enumerate(vis.trunk.blocks):
                h = blk(h)
                if i in self.extract_layers:
                    taps.append(h)
            return h, taps
        if hasattr(vis, "transformer"):                          # OpenAI-CLIP layout: blocks take a [N, B, D] view
            h = vis.embed(x).permute(1, 0, 2)
            for i, blk in enumerate(vis.transformer.resblocks):
                h = blk(h)
                if i in self.extract_layers:
                    taps.append(h.permute(1, 0, 2))
            return h.permute(1, 0, 2), taps
        raise AttributeError("Model visual encoder has neither 'trunk' nor 'transformer' attribute")

    def extract_vit_features(self, x):
        return self._tower_taps(x)

    def forward(self, x):
        _, taps = self.extract_vit_features(x)
        fused = None
        for lvl in reversed(range(len(taps))):                   # deep to shallow, summed
            t = taps[lvl][:, 1:, :].to(self.reduces[lvl].weight.dtype)   # drop the CLS token
            y = self.blocks[lvl](self.reduces[lvl](t))
            fused = y if fused is None else fused + y
        B, n_tok, _ = fused.shape
        g = int(math.sqrt(n_tok))
        fmap = fused.transpose(1, 2).reshape(B, self.reduce_dim, g, g)
        return self.seg_head(fmap) if self.task == "seg" else self.cls_head(fmap)

    def freeze_clip_backbone(self):
        """Backbone frozen except parameters whose name contains adapter / mona / lora; head modules trainable."""
        for name, p in self.clip_model.named_parameters():
            p.requires_grad = any(k in name for k in ("adapter", "mona", "lora"))
        head = self.seg_head if self.task == "seg" else self.cls_head
        for mod in (self.reduces, self.blocks, head):
            for p in mod.parameters():
                p.requires_grad = True


class TimmCLIPAdapter(_PyramidHeadBase):
    """reference src/third_party/timm/clip_adapter.py (BiomedCLIP / UniMedCLIP trunks, MetaCLIP transformer)."""

    def __init__(self, clip_model, extract_layers=[3, 6, 9], reduce_dim=512, num_classes=2, img_size=224, patch_size=16, task="seg"):
        vis = clip_model.visual
        dim = vis.trunk.embed_dim if hasattr(vis, "trunk") else vis.transformer.width
        head = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Dropout(0.5), nn.Linear(reduce_dim, num_classes))
        super().__init__(clip_model, dim, extract_layers, reduce_dim, num_classes, img_size, task, head)


class CLIPAdapter(_PyramidHeadBase):
    """reference src/third_party/openai_clip/clip_adapter.py (OpenAI CLIP towers)."""

    def __init__(self, clip_model, extract_layers=[3, 6, 9], reduce_dim=512, num_classes=2, img_size=224, patch_size=16, task="seg"):
        head = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(reduce_dim, reduce_dim), nn.ReLU(), nn.Dropout(0.1),
                             nn.Linear(reduce_dim, num_classes))
        super().__init__(clip_model, clip_model.visual.transformer.width, extract_layers, reduce_dim, num_classes, img_size, task, head)
