"""InfoNCE on the B200 kernels — host-side mirror of the reference's src/losses/losses.py:10-47.

Same class name, constructor (`temperature=0.07`) and call signature
`forward(image_features, text_features, batch_size=None) -> 0-d tensor`.  Forward and backward are
produced by one fused kernel sequence (normalise -> logits / LSEs / loss / d-features); the only
autograd node is the scalar scale in backward.

Data parallel (added; the reference is single-GPU): with `gather_distributed=True` and an initialised
torch.distributed group, the L2-normalised features of all ranks are all-gathered (NCCL over NVLink),
every rank evaluates the global [Bg,Bg] logit matrix and keeps the gradient rows of its own slice —
exactly the reference loss applied to the concatenated global batch.
"""
import os

import torch
import torch.nn as nn

from . import ops


class _InfoNCEFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, txt, temperature, gather):
        import torch.distributed as dist
        world, rank = 1, 0
        if gather and dist.is_available() and dist.is_initialized():
            world, rank = dist.get_world_size(), dist.get_rank()
        img, txt = img.contiguous(), txt.contiguous()
        Bl, E = img.shape
        # Data-parallel contract: every rank passes the SAME local batch size (the global batch is Bl * world and the gather is
        # all_gather_into_tensor), i.e. the sampler drops a ragged final batch (DistributedSampler(drop_last=True)) -- the
        # reference's single-process loop has no such constraint (finetune.py:245-300).  NGU_DEBUG_BATCH=1 verifies it with one
        # extra collective per step instead of hanging in the gather.
        if world > 1 and os.environ.get("NGU_DEBUG_BATCH") == "1":
            sizes = torch.tensor([Bl, -Bl], device=img.device, dtype=torch.int64)
            dist.all_reduce(sizes, op=dist.ReduceOp.MAX)
            if int(sizes[0]) != Bl or int(-sizes[1]) != Bl:
                raise ValueError(f"InfoNCE under data parallelism needs equal local batch sizes on every rank (this rank: {Bl}, "
                                 f"range over ranks: {int(-sizes[1])}..{int(sizes[0])}); use drop_last in the sampler")
        Bg, r0 = Bl * world, Bl * rank
        dev = img.device
        ihat = torch.empty(Bg, E, device=dev, dtype=torch.float32)
        that = torch.empty(Bg, E, device=dev, dtype=torch.float32)
        ni = torch.empty(Bl, device=dev, dtype=torch.float32)
        nt = torch.empty(Bl, device=dev, dtype=torch.float32)
        ops.infonce_normalize(img, ihat[r0:r0 + Bl], ni)
        ops.infonce_normalize(txt, that[r0:r0 + Bl], nt)
        if world > 1:
            # one collective for both modalities: stack local slices, gather, unstack
            loc = torch.stack([ihat[r0:r0 + Bl], that[r0:r0 + Bl]], 0)           # [2, Bl, E]
            allb = torch.empty(world, 2, Bl, E, device=dev, dtype=torch.float32)
            from . import _segcap
            _segcap.collective(lambda: dist.all_gather_into_tensor(allb, loc))   # a segment boundary under graph capture
            ihat = allb[:, 0].reshape(Bg, E).contiguous()
            that = allb[:, 1].reshape(Bg, E).contiguous()
        want = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        # bf16 features: logits / feature gradients on the tcgen05 GEMM; fp32 features (check mode): CUDA cores
        tc = img.dtype == torch.bfloat16 and txt.dtype == torch.bfloat16
        loss, di, dt_ = ops.infonce_core(ihat, that, r0, Bl, temperature, want_grad=want, tensor_cores=tc)
        if want:
            ctx.save_for_backward(di, dt_, ihat[r0:r0 + Bl], that[r0:r0 + Bl], ni, nt)
        ctx.dtypes = (img.dtype, txt.dtype)
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        di, dt_, ih, th, ni, nt = ctx.saved_tensors
        gs = g.detach().float().contiguous().view(1)
        dI = ops.infonce_normalize_bwd(di, ih.contiguous(), ni, gs, ctx.dtypes[0]) if ctx.needs_input_grad[0] else None
        dT = ops.infonce_normalize_bwd(dt_, th.contiguous(), nt, gs, ctx.dtypes[1]) if ctx.needs_input_grad[1] else None
        return dI, dT, None, None


class InfoNCELoss(nn.Module):
    def __init__(self, temperature=0.07, gather_distributed=False):
        super(InfoNCELoss, self).__init__()
        self.temperature = temperature
        self.gather_distributed = gather_distributed

    def forward(self, image_features, text_features, batch_size=None):
        if batch_size is not None and batch_size != image_features.shape[0]:
            # the reference only uses batch_size for arange(); a mismatch would fail in F.cross_entropy there
            raise ValueError(f"batch_size {batch_size} does not match features {tuple(image_features.shape)}")
        return _InfoNCEFunction.apply(image_features, text_features, float(self.temperature), self.gather_distributed)
