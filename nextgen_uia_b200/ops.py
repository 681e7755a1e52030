"""Tensor-level wrappers over the C ABI (include/ngu_b200.h).

PyTorch is used here only for device memory and streams: every function takes CUDA tensors,
allocates outputs with torch.empty and enqueues the sm_100a kernels of libngu_b200.so on the
current stream.  dtype selects the path: torch.bfloat16 -> tcgen05 product path, torch.float32 ->
fp32 check mode (CUDA cores).  Nothing here falls back to torch math.
"""
import ctypes

import torch

from . import _lib as L

_byref = ctypes.byref


def _dt(t):
    if t.dtype == torch.bfloat16:
        return L.NGU_BF16
    if t.dtype == torch.float32:
        return L.NGU_F32
    raise L.NguError(f"unsupported activation dtype {t.dtype} (bf16 or fp32)")


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise L.NguError("ngu ops need CUDA tensors: there is no CPU fallback in the product path")


def _f32(t):
    assert t.dtype == torch.float32 and t.is_contiguous(), "parameters are passed as contiguous fp32"
    return t


# ------------------------------------------------------------------------------------------------
def gemm(A, B, *, bias=None, act=L.ACT_NONE, aux=None, aux_mode=L.AUX_NONE, save_pre=False,
         A2=None, B2=None, alpha=1.0, out=None, block_n=0, force_simt=False, aux2=None, rowab=None):
    """C[M,N] = epi(alpha * (A[M,K] @ B[N,K]^T + A2 @ B2^T)); returns C, or (C, D) with save_pre where
    D = act'(pre-activation) (the factor AUX_DACT multiplies by in backward; the pre-activation itself if act is NONE)."""
    _need_cuda(A, B)
    assert A.dim() == 2 and B.dim() == 2 and A.shape[1] == B.shape[1] and A.dtype == B.dtype
    assert A.stride(1) == 1 and B.stride(1) == 1
    M, K = A.shape
    N = B.shape[0]
    C = out if out is not None else torch.empty(M, N, device=A.device, dtype=A.dtype)
    assert C.shape == (M, N) and C.stride(1) == 1 and C.dtype == A.dtype
    # bf16 path with an activation: the saved derivative act'(pre) travels as ONE byte per element (include/ngu_b200.h save_pre == 2)
    pre_u8 = bool(save_pre) and A.dtype == torch.bfloat16 and act != L.ACT_NONE and N % 16 == 0 and aux is None and not force_simt
    Pre = torch.empty(M, N, device=A.device, dtype=(torch.uint8 if pre_u8 else A.dtype)) if save_pre else None
    d = L.GemmDesc()
    d.A, d.lda, d.B, d.ldb, d.C, d.ldc = A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), C.data_ptr(), C.stride(0)
    if A2 is not None:
        assert A2.shape[0] == M and B2.shape[0] == N and A2.shape[1] == B2.shape[1]
        assert A2.stride(1) == 1 and B2.stride(1) == 1 and A2.dtype == A.dtype and B2.dtype == A.dtype
        d.A2, d.lda2, d.B2, d.ldb2, d.K2 = A2.data_ptr(), A2.stride(0), B2.data_ptr(), B2.stride(0), A2.shape[1]
    if bias is not None:
        d.bias = _f32(bias).data_ptr()
    if aux is not None:
        if aux.dtype == torch.uint8:
            assert aux_mode == L.AUX_DACT and A.dtype == torch.bfloat16, "one-byte operand = saved activation derivative of the bf16 path"
            aux_mode = L.AUX_DACT_U8
        else:
            assert aux.dtype == A.dtype
        assert aux.shape == (M, N) and aux.stride(1) == 1
        d.aux, d.ldaux = aux.data_ptr(), aux.stride(0)
    if save_pre:
        d.Pre, d.ldpre = Pre.data_ptr(), Pre.stride(0)
    if aux2 is not None:
        assert aux2.shape == (M, N) and aux2.stride(1) == 1 and aux2.dtype == A.dtype
        d.aux2, d.ldaux2 = aux2.data_ptr(), aux2.stride(0)
    if rowab is not None:
        assert rowab.shape == (M, 2) and rowab.dtype == torch.float32 and rowab.is_contiguous()
        d.rowab = rowab.data_ptr()
    d.M, d.N, d.K = M, N, K
    d.act, d.aux_mode, d.save_pre = act, aux_mode, (2 if pre_u8 else int(save_pre))
    d.alpha = alpha
    d.dtype = (100 + L.NGU_BF16) if (force_simt and A.dtype == torch.bfloat16) else _dt(A)
    d.block_n = block_n
    L.check(L.lib().ngu_gemm(_byref(d), _stream()), "ngu_gemm")
    return (C, Pre) if save_pre else C


def ln_fwd(x, w, b, eps, *, gamma=None, gammax=None, rows=None, ldx=None, out=None, ldy=None, save_stats=True):
    """LayerNorm over the last dim of x (viewed as [M, D]); optional Mona pre-scale.  Returns (y, mean, rstd)."""
    _need_cuda(x)
    D = x.shape[-1]
    if rows is None:
        assert x.is_contiguous()
        M, ldx_ = x.numel() // D, D
    else:
        M, ldx_ = rows, ldx
    y = out if out is not None else torch.empty((M, D) if rows is not None else x.shape, device=x.device, dtype=x.dtype)
    mean = torch.empty(M, device=x.device, dtype=torch.float32) if save_stats else None
    rstd = torch.empty(M, device=x.device, dtype=torch.float32) if save_stats else None
    d = L.LnDesc()
    d.x, d.ldx, d.y, d.ldy = x.data_ptr(), ldx_, y.data_ptr(), (ldy if ldy is not None else D)
    d.w, d.b = _f32(w).data_ptr(), _f32(b).data_ptr()
    d.gamma, d.gammax = _p(gamma), _p(gammax)
    d.mean, d.rstd = _p(mean), _p(rstd)
    d.M, d.D, d.eps, d.dtype = M, D, eps, _dt(x)
    L.check(L.lib().ngu_ln_fwd(_byref(d), _stream()), "ngu_ln_fwd")
    return y, mean, rstd


def ln_bwd(g, x, mean, rstd, w, *, dres=None, rows=None, ldg=None, ldx=None, out=None, lddx=None):
    """dx = LNbwd(g) (+ dres) for a frozen-affine LayerNorm."""
    _need_cuda(g, x)
    D = x.shape[-1]
    M = rows if rows is not None else x.numel() // D
    dx = out if out is not None else torch.empty((M, D) if rows is not None else x.shape, device=x.device, dtype=x.dtype)
    d = L.LnBwdDesc()
    d.g, d.ldg = g.data_ptr(), (ldg if ldg is not None else D)
    d.x, d.ldx = x.data_ptr(), (ldx if ldx is not None else D)
    if dres is not None:
        assert dres.is_contiguous()
        d.dres, d.ldr = dres.data_ptr(), D
    d.dx, d.lddx = dx.data_ptr(), (lddx if lddx is not None else D)
    d.mean, d.rstd, d.w = mean.data_ptr(), rstd.data_ptr(), _f32(w).data_ptr()
    d.M, d.D, d.dtype = M, D, _dt(x)
    L.check(L.lib().ngu_ln_bwd(_byref(d), _stream()), "ngu_ln_bwd")
    return dx


def mona_pre_bwd(du, dy, x, mean, rstd, w, b, gamma, gammax, dw, db, dgamma, dgammax, dycol):
    _need_cuda(du, dy, x)
    D = x.shape[-1]
    M = x.numel() // D
    dx = torch.empty_like(x)
    d = L.MonaPreBwdDesc()
    d.du, d.dy, d.x = du.data_ptr(), dy.data_ptr(), x.data_ptr()
    d.mean, d.rstd = mean.data_ptr(), rstd.data_ptr()
    d.w, d.b, d.gamma, d.gammax = (_f32(t).data_ptr() for t in (w, b, gamma, gammax))
    d.dx = dx.data_ptr()
    d.dw, d.db, d.dgamma, d.dgammax, d.dycol = (_f32(t).data_ptr() for t in (dw, db, dgamma, dgammax, dycol))
    d.M, d.D, d.dtype = M, D, _dt(x)
    L.check(L.lib().ngu_mona_pre_bwd(_byref(d), _stream()), "ngu_mona_pre_bwd")
    return dx


def _conv_desc(h, weights, hw, has_cls, drop_p, seed):
    B, N, C = h.shape
    d = L.MonaConvDesc()
    k3, b3, k5, b5, k7, b7, P, bp = weights[:8]
    d.w.k3, d.w.b3, d.w.k5, d.w.b5, d.w.k7, d.w.b7, d.w.P, d.w.bp = (_f32(t).data_ptr() for t in (k3, b3, k5, b5, k7, b7, P, bp))
    if len(weights) > 8:  # variants: (freq, ne_w1, ne_b1, ne_w2, ne_b2), each tensor or None
        freq, w1, b1, w2, b2 = weights[8:13]
        d.w.freq = _p(freq)
        d.w.ne_w1, d.w.ne_b1, d.w.ne_w2, d.w.ne_b2 = _p(w1), _p(b1), _p(w2), _p(b2)
    d.B, d.N, d.H, d.W, d.C, d.has_cls = B, N, hw[0], hw[1], C, int(has_cls)
    d.drop_p, d.seed, d.dtype = float(drop_p), int(seed) & 0xFFFFFFFFFFFFFFFF, _dt(h)
    return d


def mona_conv_fwd(h, weights, hw, has_cls, drop_p=0.0, seed=0):
    """h [B,N,C] -> g = dropout(gelu(conv-stage(h))); weights = (k3,b3,k5,b5,k7,b7,P,bp[,freq,ne_w1,ne_b1,ne_w2,ne_b2]) fp32."""
    _need_cuda(h)
    assert h.is_contiguous()
    g = torch.empty_like(h)
    d = _conv_desc(h, weights, hw, has_cls, drop_p, seed)
    d.h, d.g = h.data_ptr(), g.data_ptr()
    L.check(L.lib().ngu_mona_conv_fwd(_byref(d), _stream()), "ngu_mona_conv_fwd")
    return g


def mona_conv_bwd(h, dg, weights, grads, hw, has_cls, drop_p=0.0, seed=0):
    """Returns dh; accumulates into grads = (dk3,db3,dk5,db5,dk7,db7,dP,dbp,db1) fp32."""
    _need_cuda(h, dg)
    assert h.is_contiguous() and dg.is_contiguous()
    dh = torch.empty_like(h)
    d = _conv_desc(h, weights, hw, has_cls, drop_p, seed)
    d.h, d.dg, d.dh = h.data_ptr(), dg.data_ptr(), dh.data_ptr()
    (d.gr.dk3, d.gr.db3, d.gr.dk5, d.gr.db5, d.gr.dk7, d.gr.db7, d.gr.dP, d.gr.dbp, d.gr.db1) = (_f32(t).data_ptr() for t in grads[:9])
    if len(grads) > 9:
        d.gr.dfreq, d.gr.dne_w1, d.gr.dne_b1, d.gr.dne_w2, d.gr.dne_b2 = (_p(t) for t in grads[9:14])
    L.check(L.lib().ngu_mona_conv_bwd(_byref(d), _stream()), "ngu_mona_conv_bwd")
    return dh


# ------------------------------------------------------------------------------------------------
# fused Mona path (bf16): derived operands + stage kernels (include/ngu_b200.h "Fused Mona adapter")
class MonaDerivedBuffers:
    """Device buffers of `ngu_mona_derived` for one adapter, carved out of one allocation."""

    def __init__(self, D, device):
        sizes = [("wab", 128 * D * 2), ("wcat_t", D * 128 * 2), ("w2", D * 64 * 2), ("w2_t", 64 * D * 2), ("ca", 256), ("cb", 256),
                 ("kc", 49 * 64 * 4), ("bc", 256), ("pb", 64 * 64 * 2), ("bp", 256)]
        total = sum((n + 255) // 256 * 256 for _, n in sizes)
        self.buf = torch.empty(total, device=device, dtype=torch.uint8)
        self.D = D
        self.c = L.MonaDerived()
        off = 0
        base = self.buf.data_ptr()
        self.off = {}
        for name, n in sizes:
            setattr(self.c, name, base + off)
            self.off[name] = (off, n)
            off += (n + 255) // 256 * 256

    def view(self, name, dtype, shape):
        off, n = self.off[name]
        return self.buf[off:off + n].view(dtype).view(*shape)


def mona_params_struct(m):
    """ngu_mona_params of a BaselineMona / FreqEnhancedMona module (fp32 parameters, reference shapes)."""
    c = m.adapter_conv
    P = L.MonaParams()
    P.w1, P.b1, P.w2, P.b2 = (_f32(t.detach()).data_ptr() for t in (m.project1.weight, m.project1.bias, m.project2.weight, m.project2.bias))
    P.ln_w, P.ln_b, P.gamma, P.gammax = (_f32(t.detach()).data_ptr() for t in (m.norm.weight, m.norm.bias, m.gamma, m.gammax))
    cw = P.conv
    cw.k3, cw.b3, cw.k5, cw.b5, cw.k7, cw.b7 = (_f32(t.detach()).data_ptr() for t in (c.conv1.weight, c.conv1.bias, c.conv2.weight, c.conv2.bias,
                                                                                       c.conv3.weight, c.conv3.bias))
    cw.P, cw.bp = _f32(c.projector.weight.detach()).data_ptr(), _f32(c.projector.bias.detach()).data_ptr()
    freq = getattr(c, "freq_filter", None)
    cw.freq = _p(freq.detach()) if freq is not None else None
    return P


class MonaPrepPlan:
    """One-launch refresh (`ngu_mona_prep`) of the derived operands of a set of adapters after each optimiser update."""

    def __init__(self, monas):
        import numpy as np
        self.monas = list(monas)
        dev = self.monas[0].project1.weight.device
        self.D = self.monas[0].project1.weight.shape[1]
        items = (L.MonaPrepItem * len(self.monas))()
        for i, m in enumerate(self.monas):
            if getattr(m, "_ngu_derived", None) is None or m._ngu_derived.buf.device != dev:
                m._ngu_derived = MonaDerivedBuffers(self.D, dev)
            items[i].p = mona_params_struct(m)
            items[i].d = m._ngu_derived.c
        raw = np.frombuffer(bytes(items), dtype=np.uint8).copy()
        self.table = torch.from_numpy(raw).to(dev)
        self.ptrs = [p.data_ptr() for m in self.monas for p in m.parameters()]

    def valid(self):
        return self.ptrs == [p.data_ptr() for m in self.monas for p in m.parameters()]

    def run(self):
        L.check(L.lib().ngu_mona_prep(self.table.data_ptr(), len(self.monas), self.D, _stream()), "ngu_mona_prep")
        for m in self.monas:
            m._ngu_derived_key = _mona_param_key(m)


def _mona_param_key(m):
    return (PARAM_EPOCH[0],) + tuple((p.data_ptr(), p._version) for p in m.parameters())


def mona_derived(m):
    """Fresh derived operands of adapter `m` (re-made when a parameter changed since the last prep)."""
    if getattr(m, "_ngu_derived_key", None) != _mona_param_key(m) or getattr(m, "_ngu_derived", None) is None:
        plan = getattr(m, "_ngu_own_plan", None)
        if plan is None or not plan.valid():
            plan = m._ngu_own_plan = MonaPrepPlan([m])
        plan.run()
    return m._ngu_derived


_MONA_WS = {}


def mona_ws(D, device):
    key = (D, str(device))
    if key not in _MONA_WS:
        _MONA_WS[key] = torch.zeros(int(L.lib().ngu_mona_ws_floats(D)), device=device, dtype=torch.float32)
    return _MONA_WS[key]


def _mona_stage_desc(der, B, N, D, hw, has_cls, drop_p, seed, eps):
    d = L.MonaStageDesc()
    d.d = der.c
    d.B, d.N, d.H, d.W, d.D, d.has_cls = B, N, hw[0], hw[1], D, int(has_cls)
    d.eps, d.drop_p, d.seed = float(eps), float(drop_p), int(seed) & 0xFFFFFFFFFFFFFFFF
    return d


def mona_fwd_stage(x, der, hw, has_cls, drop_p, seed, eps):
    """x [B,N,D] bf16 -> (h, hA, g [B,N,64] bf16, mean, rstd [B*N] fp32)."""
    _need_cuda(x)
    assert x.is_contiguous() and x.dtype == torch.bfloat16
    B, N, D = x.shape
    h = torch.empty(B, N, 64, device=x.device, dtype=x.dtype)
    hA, g = torch.empty_like(h), torch.empty_like(h)
    mean = torch.empty(B * N, device=x.device, dtype=torch.float32)
    rstd = torch.empty_like(mean)
    d = _mona_stage_desc(der, B, N, D, hw, has_cls, drop_p, seed, eps)
    d.x, d.h, d.hA, d.g, d.mean, d.rstd = x.data_ptr(), h.data_ptr(), hA.data_ptr(), g.data_ptr(), mean.data_ptr(), rstd.data_ptr()
    L.check(L.lib().ngu_mona_fwd_stage(_byref(d), _stream()), "ngu_mona_fwd_stage")
    return h, hA, g, mean, rstd


def mona_bwd_stage(h, hA, dg, mean, rstd, der, D, hw, has_cls, drop_p, seed, dP, dbp):
    """-> (dhcat [B*N,128] bf16, rowab [B*N,2] fp32, ws); dP / dbp accumulate."""
    B, N, _ = h.shape
    dhcat = torch.empty(B * N, 128, device=h.device, dtype=h.dtype)
    rowab = torch.empty(B * N, 2, device=h.device, dtype=torch.float32)
    ws = mona_ws(D, h.device)
    d = _mona_stage_desc(der, B, N, D, hw, has_cls, drop_p, seed, 0.0)
    d.h, d.hA, d.dg, d.mean, d.rstd = h.data_ptr(), hA.data_ptr(), dg.data_ptr(), mean.data_ptr(), rstd.data_ptr()
    d.dhcat, d.rowab, d.ws, d.dP, d.dbp = dhcat.data_ptr(), rowab.data_ptr(), ws.data_ptr(), _f32(dP).data_ptr(), _f32(dbp).data_ptr()
    L.check(L.lib().ngu_mona_bwd_stage(_byref(d), _stream()), "ngu_mona_bwd_stage")
    return dhcat, rowab, ws


def mona_finish(m, grads, ws, D):
    """grads: dict name -> fp32 tensor (dw1, db1, dln_w, dln_b, dgamma, dgammax, dk3, db3, dk5, db5, dk7, db7[, dfreq])."""
    P = mona_params_struct(m)
    G = L.MonaGrads()
    for k, t in grads.items():
        setattr(G, k, _f32(t).data_ptr() if t is not None else None)
    L.check(L.lib().ngu_mona_finish(_byref(P), _byref(G), ws.data_ptr(), D, _stream()), "ngu_mona_finish")


def _attn_desc(q, k, v, o, B, H, N, S, dh, strides, scale, causal, impl, kv_len=None):
    d = L.AttnDesc()
    (d.q_bs, d.q_ts), (d.k_bs, d.k_ts), (d.v_bs, d.v_ts), (d.o_bs, d.o_ts) = strides
    d.q, d.k, d.v, d.o = q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr()
    d.B, d.H, d.N, d.S, d.dh = B, H, N, S, dh
    d.scale, d.causal, d.dtype, d.impl = scale, int(causal), _dt(q), impl
    if kv_len is not None:
        assert kv_len.dtype == torch.int32 and kv_len.is_cuda and kv_len.numel() == B and kv_len.is_contiguous()
        d.kv_len = kv_len.data_ptr()
    return d


def attn_fwd_packed(qkv, B, N, H, dh, *, causal=False, impl=0, kv_len=None):
    """timm layout: qkv [B*N, 3*H*dh] (= [B,N,3,H,dh]); returns (o [B*N, H*dh], lse [B,H,N])."""
    _need_cuda(qkv)
    D = H * dh
    o = torch.empty(B * N, D, device=qkv.device, dtype=qkv.dtype)
    lse = torch.empty(B, H, N, device=qkv.device, dtype=torch.float32)
    st = ((N * 3 * D, 3 * D),) * 3 + ((N * D, D),)
    d = _attn_desc(qkv, qkv[:, D:], qkv[:, 2 * D:], o, B, H, N, N, dh, st, dh ** -0.5, causal, impl, kv_len)
    d.lse = lse.data_ptr()
    L.check(L.lib().ngu_attn_fwd(_byref(d), _stream()), "ngu_attn_fwd")
    return o, lse


def attn_bwd_packed(qkv, o, lse, do, B, N, H, dh, *, causal=False, impl=0, kv_len=None):
    """Returns dqkv [B*N, 3*H*dh]."""
    _need_cuda(qkv, o, do)
    D = H * dh
    assert do.is_contiguous() and o.is_contiguous()
    dqkv = torch.empty_like(qkv)
    st = ((N * 3 * D, 3 * D),) * 3 + ((N * D, D),)
    d = _attn_desc(qkv, qkv[:, D:], qkv[:, 2 * D:], o, B, H, N, N, dh, st, dh ** -0.5, causal, impl, kv_len)
    d.lse, d.d_o = lse.data_ptr(), do.data_ptr()
    d.dq, d.dk, d.dv = dqkv.data_ptr(), dqkv[:, D:].data_ptr(), dqkv[:, 2 * D:].data_ptr()
    if N > 256 and qkv.dtype == torch.bfloat16 and dh == 64:
        ws = torch.empty(B * H * N, device=qkv.device, dtype=torch.float32)    # delta of the key-tiled tcgen05 backward
        d.ws = ws.data_ptr()
    L.check(L.lib().ngu_attn_bwd(_byref(d), _stream()), "ngu_attn_bwd")
    return dqkv


def attn_fwd_strided(q, k, v, B, H, N, S, dh, strides, *, causal=False, impl=0):
    """General layout (see include/ngu_b200.h): strides = ((q_bs,q_ts),(k_bs,k_ts),(v_bs,v_ts),(o_bs,o_ts)); o is
    allocated as [B*N... ] by the caller's convention: returns (o flat [B*N*H*dh], lse)."""
    _need_cuda(q, k, v)
    o = torch.empty(B * N * H * dh, device=q.device, dtype=q.dtype)
    lse = torch.empty(B, H, N, device=q.device, dtype=torch.float32)
    d = _attn_desc(q, k, v, o, B, H, N, S, dh, strides, dh ** -0.5, causal, impl)
    d.lse = lse.data_ptr()
    L.check(L.lib().ngu_attn_fwd(_byref(d), _stream()), "ngu_attn_fwd")
    return o, lse


def attn_bwd_strided(q, k, v, o, lse, do, B, H, N, S, dh, strides, *, causal=False, impl=0):
    _need_cuda(q, k, v, o, do)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    d = _attn_desc(q, k, v, o, B, H, N, S, dh, strides, dh ** -0.5, causal, impl)
    d.lse, d.d_o = lse.data_ptr(), do.data_ptr()
    d.dq, d.dk, d.dv = dq.data_ptr(), dk.data_ptr(), dv.data_ptr()
    L.check(L.lib().ngu_attn_bwd(_byref(d), _stream()), "ngu_attn_bwd")
    return dq, dk, dv


def wgrad(X, Y, out=None, impl=0):
    """D[Mo,No] (+)= X[T,Mo]^T @ Y[T,No] in fp32."""
    _need_cuda(X, Y)
    assert X.shape[0] == Y.shape[0] and X.stride(1) == 1 and Y.stride(1) == 1 and X.dtype == Y.dtype
    T, Mo = X.shape
    No = Y.shape[1]
    D = out if out is not None else torch.zeros(Mo, No, device=X.device, dtype=torch.float32)
    assert D.dtype == torch.float32 and D.stride(1) == 1
    L.check(L.lib().ngu_wgrad(X.data_ptr(), X.stride(0), Y.data_ptr(), Y.stride(0), D.data_ptr(), D.stride(0), T, Mo, No,
                              _dt(X), impl, _stream()), "ngu_wgrad")
    return D


def colsum(X, out=None):
    _need_cuda(X)
    assert X.dim() == 2 and X.stride(1) == 1
    T, C = X.shape
    o = out if out is not None else torch.zeros(C, device=X.device, dtype=torch.float32)
    L.check(L.lib().ngu_colsum(X.data_ptr(), X.stride(0), o.data_ptr(), T, C, _dt(X), _stream()), "ngu_colsum")
    return o


def dropout(x, p, seed, out=None, accumulate=False):
    _need_cuda(x)
    assert x.is_contiguous()
    o = out if out is not None else torch.empty_like(x)
    L.check(L.lib().ngu_dropout(x.data_ptr(), o.data_ptr(), x.numel(), float(p), int(seed) & 0xFFFFFFFFFFFFFFFF, int(accumulate),
                                _dt(x), _stream()), "ngu_dropout")
    return o


def patchify(img, P, dtype):
    _need_cuda(img)
    assert img.dtype == torch.float32 and img.is_contiguous() and img.dim() == 4 and img.shape[1] == 3
    B, _, R, R2 = img.shape
    assert R == R2
    G = R // P
    K = 3 * P * P
    Kp = (K + 7) // 8 * 8                     # row pitch the GEMM can consume (P = 14: 588 -> 592, tail stays zero)
    out = torch.empty(B * G * G, K, device=img.device, dtype=dtype) if Kp == K else torch.zeros(B * G * G, Kp, device=img.device, dtype=dtype)
    L.check(L.lib().ngu_patchify(img.data_ptr(), out.data_ptr(), B, R, P, _dt(out), _stream()), "ngu_patchify")
    return out


def assemble_tokens(patch, cls, pos, B):
    _need_cuda(patch)
    np_, D = patch.shape[0] // B, patch.shape[1]
    out = torch.empty(B, np_ + 1, D, device=patch.device, dtype=patch.dtype)
    L.check(L.lib().ngu_assemble_tokens(patch.data_ptr(), _f32(cls).data_ptr(), _f32(pos).data_ptr(), out.data_ptr(), B, np_, D,
                                        _dt(patch), _stream()), "ngu_assemble_tokens")
    return out


def embed_tokens(ids, word, pos, type0, dtype):
    _need_cuda(ids)
    assert ids.dtype == torch.int64 and ids.is_contiguous()
    B, S = ids.shape
    V, D = word.shape
    out = torch.empty(B * S, D, device=ids.device, dtype=dtype)
    L.check(L.lib().ngu_embed_tokens(ids.data_ptr(), _f32(word).data_ptr(), _f32(pos).data_ptr(), _f32(type0).data_ptr(),
                                     out.data_ptr(), B, S, D, V, _dt(out), _stream()), "ngu_embed_tokens")
    return out


# Bumped whenever parameters are rewritten behind autograd's back (the fused optimizer updates the flat parameter buffer
# through raw pointers, so tensor version counters do not move).  Low-precision shadows are keyed on it.
PARAM_EPOCH = [0]


def bump_param_epoch():
    PARAM_EPOCH[0] += 1


class CastPlan:
    """One-launch refresh of the low-precision shadows of a set of fp32 parameters (`ngu_cast_f32_batch`).

    `entries`: list of (param, transpose).  Output tensors and the device-side item table are allocated once; `run()`
    re-converts everything (call it after the parameters changed), `get(param, transpose)` returns the shadow.
    Valid as long as the parameters keep their storage (they do: the optimizer updates in place)."""

    def __init__(self, entries, dtype):
        import numpy as np
        self.dtype = dtype
        self.outs = {}
        items = (L.CastItem * len(entries))()
        self._keep = []
        for i, (p, tr) in enumerate(entries):
            w = p.detach()
            _need_cuda(w)
            if w.dtype != torch.float32 or w.dim() != 2 or not w.is_contiguous():
                raise ValueError("CastPlan: parameters must be contiguous 2-D fp32 tensors")
            rows, cols = w.shape
            out = torch.empty((cols, rows) if tr else (rows, cols), device=w.device, dtype=dtype)
            items[i] = L.CastItem(w.data_ptr(), out.data_ptr(), rows, cols, int(tr), 1.0)
            self.outs[(id(p), bool(tr))] = out
            self._keep.append(p)
        raw = np.frombuffer(bytes(items), dtype=np.uint8).copy()
        self.table = torch.from_numpy(raw).to(self._keep[0].device)
        self.n = len(entries)
        self.ptrs = [p.data_ptr() for p in self._keep]
        self._fresh = False
        self._versions = None

    def valid(self):
        return all(p.data_ptr() == q for p, q in zip(self._keep, self.ptrs))

    def run(self):
        code = L.NGU_BF16 if self.dtype == torch.bfloat16 else L.NGU_F32
        L.check(L.lib().ngu_cast_f32_batch(self.table.data_ptr(), self.n, code, _stream()), "ngu_cast_f32_batch")
        self._versions = [p._version for p in self._keep]
        self._fresh = True

    @property
    def fresh(self):
        """True while the shadows match the parameters: set by run(), cleared by the optimizer (whose fused kernel writes
        the parameters through raw pointers) and by any in-place torch update (version counters) or re-allocation."""
        return (self._fresh and self.valid() and self._versions == [p._version for p in self._keep])

    @fresh.setter
    def fresh(self, v):
        self._fresh = bool(v)

    def get(self, p, transpose=False):
        return self.outs.get((id(p), bool(transpose)))


def cast(w, dtype, transpose=False, scale=1.0):
    """fp32 [rows, cols] parameter -> dtype copy (optionally transposed, scaled)."""
    _need_cuda(w)
    w = _f32(w.detach())
    rows, cols = w.shape
    out = torch.empty((cols, rows) if transpose else (rows, cols), device=w.device, dtype=dtype)
    L.check(L.lib().ngu_cast_f32(w.data_ptr(), out.data_ptr(), rows, cols, int(transpose), float(scale), _dt(out), _stream()),
            "ngu_cast_f32")
    return out


def infonce_normalize(x, xhat_out, norm_out):
    _need_cuda(x)
    assert x.is_contiguous() and xhat_out.dtype == torch.float32
    B, E = x.shape
    L.check(L.lib().ngu_infonce_normalize(x.data_ptr(), xhat_out.data_ptr(), norm_out.data_ptr(), B, E, _dt(x), _stream()),
            "ngu_infonce_normalize")


def infonce_core(ihat, that, r0, Bl, temperature, want_grad=True, tensor_cores=False):
    """ihat/that: fp32 [Bg,E] gathered normalised features. Returns (loss[1], dihat[Bl,E], dthat[Bl,E]).
    tensor_cores=True (bf16 product path): logits and feature gradients on the tcgen05 GEMM from bf16 copies."""
    _need_cuda(ihat, that)
    Bg, E = ihat.shape
    dev = ihat.device
    tc = tensor_cores and Bg % 8 == 0 and E % 8 == 0
    loss = torch.empty(1, device=dev, dtype=torch.float32)
    ws = torch.empty(2 * Bg * Bg + 2 * Bg, device=dev, dtype=torch.float32)
    di = torch.empty(Bl, E, device=dev, dtype=torch.float32) if want_grad else None
    dt_ = torch.empty(Bl, E, device=dev, dtype=torch.float32) if want_grad else None
    d = L.InfoNceDesc()
    d.ihat, d.that, d.dihat, d.dthat = ihat.data_ptr(), that.data_ptr(), _p(di), _p(dt_)
    d.loss, d.ws = loss.data_ptr(), ws.data_ptr()
    d.Bg, d.Bl, d.r0, d.E, d.temperature = Bg, Bl, r0, E, float(temperature)
    if tc:
        keep = [cast(ihat, torch.bfloat16), cast(that, torch.bfloat16)]
        d.ihat16, d.that16 = keep[0].data_ptr(), keep[1].data_ptr()
        if want_grad:
            keep += [cast(ihat, torch.bfloat16, transpose=True), cast(that, torch.bfloat16, transpose=True),
                     torch.empty(2 * Bl * Bg, device=dev, dtype=torch.bfloat16)]
            d.ihat16_t, d.that16_t, d.g_ws = keep[2].data_ptr(), keep[3].data_ptr(), keep[4].data_ptr()
    L.check(L.lib().ngu_infonce_core(_byref(d), _stream()), "ngu_infonce_core")
    return loss, di, dt_


def infonce_normalize_bwd(dxhat, xhat, norm, gscale, dtype):
    B, E = dxhat.shape
    dx = torch.empty(B, E, device=dxhat.device, dtype=dtype)
    L.check(L.lib().ngu_infonce_normalize_bwd(dxhat.data_ptr(), xhat.data_ptr(), norm.data_ptr(), _p(gscale), dx.data_ptr(), B, E,
                                              _dt(dx), _stream()), "ngu_infonce_normalize_bwd")
    return dx


def sqnorm(x, out):
    """out[0] += sum(x^2) for a flat fp32 buffer."""
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous() and out.dtype == torch.float32
    L.check(L.lib().ngu_sqnorm(x.data_ptr(), x.numel(), out.data_ptr(), _stream()), "ngu_sqnorm")


def guard_tick(state, mode, loss=None, gsq=None):
    """Device-side loop counters / non-finite guard (include/ngu_b200.h ngu_guard_tick); state: int64[4] CUDA tensor."""
    _need_cuda(state)
    assert state.dtype == torch.int64 and state.numel() >= 4
    L.check(L.lib().ngu_guard_tick(state.data_ptr(), _p(loss), _p(gsq), int(mode), _stream()), "ngu_guard_tick")


def set_seed_counter(counter):
    """counter: int64 CUDA tensor element (kept alive by the caller) mixed into every dropout seed, or None."""
    L.check(L.lib().ngu_set_seed_counter(None if counter is None else counter.data_ptr()), "ngu_set_seed_counter")


def kv_len(ids, pad_id, flag):
    """ids int64 [B,S] -> int32 [B] valid lengths; flag (int32[1]) |= 1 if a row is not right-padded."""
    _need_cuda(ids)
    assert ids.dtype == torch.int64 and ids.is_contiguous() and flag.dtype == torch.int32
    B, S = ids.shape
    out = torch.empty(B, device=ids.device, dtype=torch.int32)
    L.check(L.lib().ngu_kv_len(ids.data_ptr(), int(pad_id), out.data_ptr(), flag.data_ptr(), B, S, _stream()), "ngu_kv_len")
    return out


def adamw_step(param, grad, m, v, *, lr, betas, eps, weight_decay, step, max_norm=0.0, gsq=None, loss=None, zero_grad=True,
               state=None, lr_min=0.0, t_max=0):
    _need_cuda(param)
    for t in (param, grad, m, v):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == param.numel()
    d = L.AdamWDesc()
    d.param, d.grad, d.m, d.v, d.n = param.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), param.numel()
    d.lr, d.beta1, d.beta2, d.eps, d.weight_decay = float(lr), float(betas[0]), float(betas[1]), float(eps), float(weight_decay)
    d.step, d.max_norm = int(step), float(max_norm)
    d.gsq, d.loss, d.zero_grad = _p(gsq), _p(loss), int(zero_grad)
    d.state, d.lr_min, d.t_max = _p(state), float(lr_min), int(t_max)
    L.check(L.lib().ngu_adamw_step(_byref(d), _stream()), "ngu_adamw_step")
