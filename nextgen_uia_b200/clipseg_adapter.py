"""CLIPSeg adapter on the B200 encoder — mirror of the reference's src/third_party/openai_clip/clipseg_adapter.py:9-110.

CLIP ViT encoder (this repo's kernels, hidden states tapped after the blocks in `decoder.config.extract_layers`, NLD)
+ text conditioning via `clip_model.encode_text` + the Hugging Face `CLIPSegDecoder` (library code: it stays PyTorch,
SURVEY.md §2.1 row 7) + the 1 -> 2 channel logit expansion.  Same class name, constructor arguments, methods
(`extract_vit_features`, `forward(x, input_ids)`, `freeze_clip_backbone`) and attribute names (`clip_model`, `decoder`,
`extract_layers`).

The reference downloads `CIDAS/clipseg-rd64-refined` and keeps its `.decoder`; without network the decoder is built
from a local `CLIPSegConfig` (random init) unless one is passed in.
"""
import os

import torch
import torch.nn as nn


class CLIPSegAdapter(nn.Module):
    def __init__(self, clip_model, decoder_base_config="CIDAS/clipseg-rd64-refined", ckpt_dir="./ckpt", decoder=None):
        super().__init__()
        self.clip_model = clip_model
        self.ckpt_dir = ckpt_dir
        if decoder is None:
            from transformers import CLIPSegConfig
            from transformers.models.clipseg.modeling_clipseg import CLIPSegDecoder, CLIPSegForImageSegmentation
            try:
                os.makedirs(self.ckpt_dir, exist_ok=True)
                decoder = CLIPSegForImageSegmentation.from_pretrained(decoder_base_config, cache_dir=self.ckpt_dir).decoder
            except Exception:  # offline: same architecture, untrained weights
                decoder = CLIPSegDecoder(CLIPSegConfig())
        self.decoder = decoder
        self.extract_layers = self.decoder.config.extract_layers

    def extract_vit_features(self, x):
        """Hidden states (NLD, batch-first) after the resblocks listed in extract_layers (reference :42-71)."""
        vis = self.clip_model.visual
        h = vis.embed(x).permute(1, 0, 2)                 # NLD -> LND view, as the reference feeds the blocks
        taps = ()
        for i, block in enumerate(vis.transformer.resblocks):
            h = block(h)
            if i in self.extract_layers:
                taps = taps + (h.permute(1, 0, 2),)       # LND -> NLD
        return taps

    def forward(self, x, input_ids=None):
        B, _, H, W = x.shape
        taps = self.extract_vit_features(x)
        cond = self.clip_model.encode_text(input_ids)
        wdt = next(self.decoder.parameters()).dtype
        out = self.decoder(hidden_states=tuple(t.to(wdt) for t in taps), conditional_embeddings=cond.to(wdt))
        logits = out[0].view(B, -1, H, W)
        if logits.shape[1] == 1:
            logits = torch.cat([-logits, logits], dim=1)  # background = -foreground (reference :93-96)
        return logits

    def freeze_clip_backbone(self):
        """Freeze CLIP, train the decoder head (reference :100-110)."""
        for p in self.clip_model.parameters():
            p.requires_grad = False
        for p in self.decoder.parameters():
            p.requires_grad = True

    def unfreeze_adapters(self, keys=("mona", "lora")):
        """Re-enable injected adapter parameters after freeze_clip_backbone() (the reference's freeze would silently
        freeze them too, SURVEY.md §3.5 gap (a)); same substring rule as finetune.py:173-175."""
        n = 0
        for name, p in self.clip_model.named_parameters():
            if any(k in name.lower() for k in keys):
                p.requires_grad = True
                n += 1
        return n
