"""Import shim for the reference's src/nets/carafe.py: content-aware reassembly up-sampling (CARAFE), same constructor and
output shape; written with F.unfold instead of tensor.unfold chains.  Plain PyTorch, off the fine-tuning hot path."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class CARAFE(nn.Module):
    def __init__(self, inC, outC, kernel_size=3, up_factor=2):
        super().__init__()
        self.kernel_size, self.up_factor = kernel_size, up_factor
        self.down = nn.Conv2d(inC, inC // 4, 1)
        self.encoder = nn.Conv2d(inC // 4, up_factor ** 2 * kernel_size ** 2, kernel_size, 1, kernel_size // 2)
        self.out = nn.Conv2d(inC, outC, 1)

    def forward(self, in_tensor):
        n, c, h, w = in_tensor.shape
        k2, s = self.kernel_size ** 2, self.up_factor
        # predicted reassembly kernels: one softmax-normalised k x k kernel per OUTPUT pixel
        ker = F.pixel_shuffle(self.encoder(self.down(in_tensor)), s)                    # [n, k2, s h, s w]
        ker = torch.softmax(ker, dim=1).view(n, k2, h, s, w, s)
        # k x k neighbourhoods of every INPUT pixel
        nb = F.unfold(in_tensor, self.kernel_size, padding=self.kernel_size // 2).view(n, c, k2, h, w)
        up = torch.einsum("nckhw,nkhawb->nchawb", nb, ker).reshape(n, c, h * s, w * s)
        return self.out(up)
