"""Zero-shot prompt-ensemble classification on the device — mirror of the scoring loop of the reference's
src/models/biomedclip/zero_shot.py:176-228 (and src/models/clip/zero_shot.py, same recipe).

The reference keeps a dict class -> normalised prompt features, and per image batch computes, per class,
`(100.0 * image_features @ text_feats.T).mean(dim=1)`, stacks the classes to [B, n_classes] logits and hands them to the
metric accumulator (argmax / softmax in src/utils/tools.py:211).  The mean over prompts commutes with the dot product, so
here the class prototypes are built once (`ngu_zero_shot_prototypes`) and a batch is scored by one fused launch
(`ngu_zero_shot_score`: L2-normalise, dot with every prototype, argmax).
"""
import torch

from . import _lib as L
from . import ops


class ZeroShotScorer:
    """scorer = ZeroShotScorer(model); scorer.set_prompts({"benign": ids[P0,77], "malignant": ids[P1,77]});
    logits, pred = scorer(images)      # logits fp32 [B, n_classes] (reference order of the dict), pred int32 [B]"""

    def __init__(self, model, scale=100.0):
        self.model = model
        self.scale = float(scale)
        self.classes = []
        self.proto = None

    @torch.no_grad()
    def set_prompts(self, prompts_by_class):
        self.classes = list(prompts_by_class.keys())
        ids = torch.cat([prompts_by_class[c] for c in self.classes], 0)
        cls = torch.cat([torch.full((prompts_by_class[c].shape[0],), i, dtype=torch.int32) for i, c in enumerate(self.classes)]).to(ids.device)
        tf = self.model.encode_text(ids).contiguous()                # [P, E] on the kernels
        self.proto = self.prototypes(tf, cls, len(self.classes))
        return self.proto

    @staticmethod
    def prototypes(text_feat, class_of_prompt, n_classes):
        ops._need_cuda(text_feat)
        P, E = text_feat.shape
        proto = torch.empty(n_classes, E, device=text_feat.device, dtype=torch.float32)
        L.check(L.lib().ngu_zero_shot_prototypes(text_feat.data_ptr(), class_of_prompt.contiguous().data_ptr(), proto.data_ptr(), P, E, n_classes,
                                                 ops._dt(text_feat), ops._stream()), "ngu_zero_shot_prototypes")
        return proto

    def score_features(self, image_feat):
        ops._need_cuda(image_feat)
        image_feat = image_feat.contiguous()
        B, E = image_feat.shape
        C = self.proto.shape[0]
        logits = torch.empty(B, C, device=image_feat.device, dtype=torch.float32)
        pred = torch.empty(B, device=image_feat.device, dtype=torch.int32)
        L.check(L.lib().ngu_zero_shot_score(image_feat.data_ptr(), self.proto.data_ptr(), logits.data_ptr(), pred.data_ptr(), B, E, C, self.scale,
                                            ops._dt(image_feat), ops._stream()), "ngu_zero_shot_score")
        return logits, pred

    @torch.no_grad()
    def __call__(self, images):
        return self.score_features(self.model.encode_image(images))
