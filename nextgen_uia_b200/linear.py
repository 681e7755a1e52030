"""Projection helper shared by the ViT block, the BERT tower and the standalone LoRA modules.

A projection is y = x W^T + b (+ s * drop(x) A^T B^T for LoRA, src/adapters/lora.py:78-90) with a
FROZEN base weight: backward only ever needs dx = dy W (dgrad), never dW, so the host keeps a
[N,K] and a [K,N] low-precision copy of every frozen weight (made once) and both directions are
"NT" tcgen05 GEMMs.  The LoRA branch is evaluated in its low-rank form and accumulated into the same
TMEM tile as extra K blocks (A2/B2 operand pair of ngu_gemm) — the reference's dense B@A is never formed.
"""
import torch
import torch.nn.functional as F

from . import _lib as L
from . import ops


def frozen_copies(weight, dtype):
    """(W [N,K], W^T [K,N]) in `dtype` for a frozen fp32 parameter, cached on the parameter."""
    if weight.requires_grad:
        raise NotImplementedError(
            "ngu B200 path: base projection weights must be frozen (Mona / LoRA fine-tuning, "
            "src/models/biomedclip/finetune.py:165-197); full fine-tuning is outside this hot path")
    key = (dtype, weight.device, weight._version, weight.data_ptr())
    cache = getattr(weight, "_ngu_cache", None)
    if cache is None or cache[0] != key:
        w32 = weight.detach().float().contiguous()
        cache = (key, ops.cast(w32, dtype), ops.cast(w32, dtype, transpose=True))
        weight._ngu_cache = cache
    return cache[1], cache[2]


class Proj:
    """Per-call view of one projection's operands."""

    __slots__ = ("W", "WT", "bias", "A", "B", "scaling", "p", "r", "rp")

    def __init__(self, linear, dtype):
        self.W, self.WT = frozen_copies(linear.weight, dtype)
        self.bias = linear.bias
        self.A = self.B = None
        self.scaling, self.p, self.r, self.rp = 0.0, 0.0, 0, 0
        r = getattr(linear, "r", 0)
        if r and hasattr(linear, "w_lora_A"):
            self.A, self.B = linear.w_lora_A, linear.w_lora_B
            self.scaling = float(linear.scaling)
            self.r = r
            # bf16: rank zero-padded to 64 so the factors' gradients run on the tcgen05 wgrad kernel (N tile = 64);
            # the base GEMM only consumes the first ceil16(r) columns as its extra K block
            self.rp = 64 if (dtype == torch.bfloat16 and r <= 64) else r
            drop = getattr(linear, "dropout", None)
            self.p = float(drop.p) if (drop is not None and linear.training) else 0.0


def _lora_shadows(pj, dtype):
    """(A [rp,K], B [N,rp], A^T [K,rp], B^T [rp,N]) in `dtype`, rank zero-padded to pj.rp.  Made once per parameter
    update (keyed on the tensors' version counters and ops.PARAM_EPOCH) and shared by forward and backward, instead of
    padding and casting the factors four times per projection per step."""
    A, B = pj.A, pj.B
    key = (A._version, B._version, ops.PARAM_EPOCH[0], A.data_ptr(), B.data_ptr(), dtype, pj.rp)
    cache = getattr(A, "_ngu_shadow", None)
    if cache is None or cache[0] != key:
        r, K = A.shape
        N = B.shape[0]
        bufs = cache[1] if (cache is not None and cache[1][0].dtype == dtype and cache[1][0].shape == (pj.rp, K)
                            and cache[1][1].shape == (N, pj.rp)) else None
        if bufs is None:   # zero padding is written once; updates only touch the live rank
            bufs = (torch.zeros(pj.rp, K, device=A.device, dtype=dtype), torch.zeros(N, pj.rp, device=A.device, dtype=dtype),
                    torch.zeros(K, pj.rp, device=A.device, dtype=dtype), torch.zeros(pj.rp, N, device=A.device, dtype=dtype))
        with torch.no_grad():
            bufs[0][:r].copy_(A)
            bufs[1][:, :r].copy_(B)
            bufs[2][:, :r].copy_(A.t())
            bufs[3][:r].copy_(B.t())
        cache = (key, bufs)
        A._ngu_shadow = cache
    return cache[1]


_ONES_COL = {}


def _ones_col(pj, device):
    """fp32 [rp] unit vector e_{rp-1} (the bias that turns t's last padding column into ones), or None when the rank leaves no
    free padding column / the projection has no trainable bias."""
    if pj.bias is None or not pj.bias.requires_grad or pj.rp != 64 or pj.r >= pj.rp - 1 or _k2(pj) >= pj.rp:
        return None
    key = (str(device), pj.rp)
    v = _ONES_COL.get(key)
    if v is None:
        v = torch.zeros(pj.rp, device=device, dtype=torch.float32)
        v[pj.rp - 1] = 1.0
        _ONES_COL[key] = v
    return v


def _k2(pj):
    """columns of the low-rank factors the base GEMM reads as its extra K block (multiple of 16 for UMMA_K)."""
    return min(pj.rp, (pj.r + 15) // 16 * 16) if pj.rp != pj.r else pj.r


def proj_fwd(x2, pj, *, act=L.ACT_NONE, aux=None, aux_mode=L.AUX_NONE, save_pre=False, seed=0):
    """Returns (out, saved) where out is y or (y, pre) and saved feeds proj_bwd."""
    bias = pj.bias.detach() if pj.bias is not None else None
    if pj.A is None:
        return ops.gemm(x2, pj.W, bias=bias, act=act, aux=aux, aux_mode=aux_mode, save_pre=save_pre), None
    dt = x2.dtype
    A16, B16, _, _ = _lora_shadows(pj, dt)
    xd = ops.dropout(x2, pj.p, seed) if pj.p > 0 else x2
    # [M, rp] = s * drop(x) A^T.  With a trainable bias the LAST padding column of t is set to 1 (a bias on the zero row of A), so
    # the weight-gradient pass dy^T t of the backward also yields the bias gradient sum_rows dy in that column: no column-sum pass
    t = ops.gemm(xd, A16, alpha=pj.scaling, bias=_ones_col(pj, x2.device))
    k2 = _k2(pj)
    out = ops.gemm(x2, pj.W, bias=bias, act=act, aux=aux, aux_mode=aux_mode, save_pre=save_pre,
                   A2=t[:, :k2], B2=B16[:, :k2])
    return out, (xd, t, seed)


def proj_bwd(dy2, pj, saved, *, need_dx=True, need_bias=False):
    """dy2 [M,N] -> (dx [M,K] or None, dbias, dA, dB); grads are fp32 in the parameter shapes."""
    want_bias = need_bias and pj.bias is not None
    if pj.A is None:
        return (ops.gemm(dy2, pj.WT) if need_dx else None), (ops.colsum(dy2) if want_bias else None), None, None
    dt = dy2.dtype
    xd, t, seed = saved
    _, _, At, Bt = _lora_shadows(pj, dt)
    dts = ops.gemm(dy2, Bt, alpha=pj.scaling)                                   # [M, rp] = s * dy B
    wg = ops.wgrad(dy2, t)                                                      # [N, rp] = dy^T t
    dB = wg[:, :pj.r].contiguous()                                              # [N, r]
    if want_bias:
        dbias = wg[:, pj.rp - 1].contiguous() if _ones_col(pj, dy2.device) is not None else ops.colsum(dy2)
    else:
        dbias = None
    dA = ops.wgrad(xd, dts)[:, :pj.r].t().contiguous()                          # [r, K]
    dx = None
    if need_dx:
        if pj.p > 0:
            dx = ops.gemm(dy2, pj.WT)
            ops.dropout(ops.gemm(dts, At), pj.p, seed, out=dx, accumulate=True)
        else:
            k2 = _k2(pj)
            dx = ops.gemm(dy2, pj.WT, A2=dts[:, :k2], B2=At[:, :k2])
    return dx, dbias, dA, dB


_CAUSAL_OK = {}


def require_causal_mask(mask, L):
    """The kernels on this path implement exactly one attention mask: CLIP's causal text mask (reference model.py:344-350,
    an [L, L] float tensor with -inf above the diagonal and 0 elsewhere).  Any other mask (additive biases, boolean or padding
    masks, Lq != S) would be silently mis-applied, so it is refused.  The content check costs one host sync and is cached per
    (storage, version)."""
    key = (mask.data_ptr(), mask._version, tuple(mask.shape), L)
    ok = _CAUSAL_OK.get(key)
    if ok is None:
        ok = False
        if mask.dim() == 2 and mask.shape[0] == L and mask.shape[1] == L and mask.is_floating_point():
            m = mask.detach().float()
            upper = torch.triu(torch.ones(L, L, dtype=torch.bool, device=m.device), 1)
            ok = bool((torch.isneginf(m) == upper).all()) and bool((m.masked_fill(upper, 0.0) == 0).all())
        if len(_CAUSAL_OK) > 64:
            _CAUSAL_OK.clear()
        _CAUSAL_OK[key] = ok
    if not ok:
        raise NotImplementedError("attn_mask must be the causal mask triu(-inf, 1) of shape [L, L] (reference model.py:344-350); "
                                  "other masks are not implemented on this path")
