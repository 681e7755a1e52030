"""Import shim for the reference's src/nets/masker.py (`Masker`, `gumbel_softmax`, the small conv helpers
and `CoordConv` stay importable with identical signatures; nothing on the fine-tuning hot path calls them).  Plain PyTorch."""
import torch
import torch.nn as nn


def conv_layer(in_dim, out_dim, kernel_size=1, padding=0, stride=1):
    """Conv2d (no bias) -> BatchNorm2d -> ReLU (reference masker.py:6-9)."""
    conv = nn.Conv2d(in_dim, out_dim, kernel_size, stride, padding, bias=False)
    return nn.Sequential(conv, nn.BatchNorm2d(out_dim), nn.ReLU(True))


def linear_layer(in_dim, out_dim, bias=False):
    """Linear -> BatchNorm1d -> ReLU (reference masker.py:12-13)."""
    return nn.Sequential(nn.Linear(in_dim, out_dim, bias), nn.BatchNorm1d(out_dim), nn.ReLU(True))


class CoordConv(nn.Module):
    """Appends normalised (x, y) coordinate planes in [-1, 1] before a conv_layer (reference masker.py:16-35)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, padding=1, stride=1):
        super().__init__()
        self.conv1 = conv_layer(in_channels + 2, out_channels, kernel_size, padding, stride)

    def add_coord(self, input):
        b, _, h, w = input.shape
        xs = torch.linspace(-1, 1, w, device=input.device).view(1, 1, 1, w).expand(b, 1, h, w)
        ys = torch.linspace(-1, 1, h, device=input.device).view(1, 1, h, 1).expand(b, 1, h, w)
        return torch.cat([input, xs, ys], dim=1)

    def forward(self, x):
        return self.conv1(self.add_coord(x))


def gumbel_softmax(logits, tau=1e-5):
    """softmax((logits + Gumbel noise) / tau) over the last dim (reference masker.py:38-41)."""
    u = torch.rand_like(logits)
    noise = -torch.log(1e-10 - torch.log(u + 1e-10))
    return torch.softmax((logits + noise) / tau, dim=-1)


class Masker(nn.Module):
    """CoordConv -> conv_layer -> CoordConv -> 3x3 conv -> BatchNorm -> sigmoid soft mask (reference masker.py:44-66)."""

    def __init__(self, in_dim=512, outdim=512):
        super().__init__()
        self.in_dim, self.outdim = in_dim, outdim
        self.conv = nn.Conv2d(in_dim, outdim, kernel_size=3, stride=1, padding=1)
        self.bn = nn.BatchNorm2d(outdim)
        self.relu = nn.ReLU()
        self.sig = nn.Sigmoid()
        self.coordconv = CoordConv(in_dim, in_dim, 3, 1)
        self.coordconv2 = CoordConv(in_dim, in_dim, 3, 1)
        self.conv2 = conv_layer(in_dim, outdim, 3, 1, 1)

    def forward(self, x):
        y = self.coordconv2(self.conv2(self.coordconv(x)))
        return self.sig(self.bn(self.conv(y)))
