"""Import shim for the reference's src/nets/layers.py: `conv_layer`, `linear_layer`, `CoordConv`,
`Projector`, `FPN_AD` with the reference's signatures and sub-module names.  Plain PyTorch; dead code in the reference too."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .carafe import CARAFE
from .masker import CoordConv, Masker, conv_layer, linear_layer  # noqa: F401  (re-exported like the reference module)


class Projector(nn.Module):
    """Up-samples a visual map 16x and correlates it with a per-sample dynamic k x k kernel predicted from a text vector."""

    def __init__(self, word_dim=1024, in_dim=256, kernel_size=3):
        super().__init__()
        self.in_dim, self.kernel_size = in_dim, kernel_size
        self.vis = nn.Sequential(nn.Upsample(scale_factor=4, mode="bilinear"), conv_layer(in_dim * 2, in_dim * 2, 3, padding=1),
                                 nn.Upsample(scale_factor=4, mode="bilinear"), conv_layer(in_dim * 2, in_dim, 3, padding=1),
                                 nn.Conv2d(in_dim, in_dim, 1))
        self.txt = nn.Linear(word_dim, in_dim * kernel_size * kernel_size + 1)

    def forward(self, x, word):
        feat = self.vis(x)
        b, c, h, w = feat.shape
        dyn = self.txt(word)
        weight = dyn[:, :-1].reshape(b, c, self.kernel_size, self.kernel_size)
        out = F.conv2d(feat.reshape(1, b * c, h, w), weight, bias=dyn[:, -1], padding=self.kernel_size // 2, groups=b)
        return out.transpose(0, 1)


class FPN_AD(nn.Module):
    """Three-level fusion that splits every level into a masked / complementary pair (Masker) and aggregates both."""

    def __init__(self, in_channels=[512, 1024, 1024], out_channels=[256, 512, 1024]):
        super().__init__()
        i0, i1, i2 = in_channels
        o0, o1, o2 = out_channels
        self.txt_proj = linear_layer(i2, o2)
        self.f1_v_proj = conv_layer(i2, o2, 1, 0)
        self.norm_layer = nn.Sequential(nn.BatchNorm2d(o2), nn.ReLU(True))
        self.f2_v_proj = conv_layer(i1, o1, 3, 1)
        self.f2_cat = conv_layer(o2 + o1, o1, 1, 0)
        self.f3_v_proj = conv_layer(i0, o0, 3, 1)
        self.f3_cat = conv_layer(o0 + o1, o1, 1, 0)
        self.f4_proj5 = conv_layer(o2, o1, 3, 1)
        self.f4_proj4 = conv_layer(o1, o1, 3, 1)
        self.f4_proj3 = conv_layer(o1, o1, 3, 1)
        self.aggr = conv_layer(3 * o1, o1, 1, 0)
        self.coordconv = nn.Sequential(CoordConv(o1, o1, 3, 1), conv_layer(o1, o1, 3, 1))
        self.masker3, self.masker4, self.masker5 = Masker(512, 512), Masker(512, 512), Masker(512, 512)
        self.carafe = CARAFE(i2, i2, up_factor=2)

    def forward(self, imgs, state):
        v3, v4, v5 = imgs
        levels = [(F.avg_pool2d(v3, 2, 2), self.masker3), (self.f4_proj4(self.f2_v_proj(v4)), self.masker4), (self.carafe(v5), self.masker5)]
        kept, dropped = [], []
        for feat, masker in levels:
            m = masker(feat)
            kept.append(feat * m)
            dropped.append(feat * (1.0 - m))
        return self.aggr(torch.cat(kept, dim=1)), self.aggr(torch.cat(dropped, dim=1))
