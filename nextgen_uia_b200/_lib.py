"""ctypes binding of libngu_b200.so (C ABI in include/ngu_b200.h).

The library is the product: there is no CPU or PyTorch fallback.  If the shared object is missing
or no sm_100 device is usable, calls raise instead of silently computing something else.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# NGU_LIB selects another in-tree build of the same sources (tools/: the -DNGU_CONV_PROF phase-timing build)
LIB_PATH = os.environ.get("NGU_LIB") or os.path.join(_HERE, "libngu_b200.so")

NGU_BF16, NGU_F32 = 0, 1
ACT_NONE, ACT_GELU, ACT_QUICKGELU = 0, 1, 2
AUX_NONE, AUX_RESIDUAL, AUX_DACT, AUX_MONA_DX, AUX_DACT_U8 = 0, 1, 2, 3, 4

_c_void_p, _c_int, _c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_float


class CastItem(ctypes.Structure):   # ngu_cast_item
    _fields_ = [("inp", _c_void_p), ("out", _c_void_p), ("rows", _c_int), ("cols", _c_int), ("transpose", _c_int),
                ("scale", _c_float)]


class GemmDesc(ctypes.Structure):
    _fields_ = [
        ("A", _c_void_p), ("lda", _c_int),
        ("B", _c_void_p), ("ldb", _c_int),
        ("C", _c_void_p), ("ldc", _c_int),
        ("A2", _c_void_p), ("lda2", _c_int),
        ("B2", _c_void_p), ("ldb2", _c_int),
        ("bias", _c_void_p),
        ("aux", _c_void_p), ("ldaux", _c_int),
        ("Pre", _c_void_p), ("ldpre", _c_int),
        ("M", _c_int), ("N", _c_int), ("K", _c_int), ("K2", _c_int),
        ("act", _c_int), ("aux_mode", _c_int), ("save_pre", _c_int),
        ("alpha", _c_float),
        ("dtype", _c_int),
        ("block_n", _c_int),
        ("aux2", _c_void_p), ("ldaux2", _c_int),
        ("rowab", _c_void_p),
        ("c_dtype", _c_int),
    ]


_i64, _u64 = ctypes.c_int64, ctypes.c_uint64


class LnDesc(ctypes.Structure):
    _fields_ = [("x", _c_void_p), ("ldx", _i64), ("y", _c_void_p), ("ldy", _i64),
                ("w", _c_void_p), ("b", _c_void_p), ("gamma", _c_void_p), ("gammax", _c_void_p),
                ("mean", _c_void_p), ("rstd", _c_void_p),
                ("M", _c_int), ("D", _c_int), ("eps", _c_float), ("dtype", _c_int)]


class LnBwdDesc(ctypes.Structure):
    _fields_ = [("g", _c_void_p), ("ldg", _i64), ("x", _c_void_p), ("ldx", _i64),
                ("dres", _c_void_p), ("ldr", _i64), ("dx", _c_void_p), ("lddx", _i64),
                ("mean", _c_void_p), ("rstd", _c_void_p), ("w", _c_void_p),
                ("M", _c_int), ("D", _c_int), ("dtype", _c_int)]


class MonaPreBwdDesc(ctypes.Structure):
    _fields_ = [("du", _c_void_p), ("dy", _c_void_p), ("x", _c_void_p),
                ("mean", _c_void_p), ("rstd", _c_void_p),
                ("w", _c_void_p), ("b", _c_void_p), ("gamma", _c_void_p), ("gammax", _c_void_p),
                ("dx", _c_void_p),
                ("dw", _c_void_p), ("db", _c_void_p), ("dgamma", _c_void_p), ("dgammax", _c_void_p), ("dycol", _c_void_p),
                ("M", _c_int), ("D", _c_int), ("dtype", _c_int)]


class MonaConvWeights(ctypes.Structure):
    _fields_ = [(n, _c_void_p) for n in ("k3", "b3", "k5", "b5", "k7", "b7", "P", "bp", "freq", "ne_w1", "ne_b1", "ne_w2", "ne_b2")]


class MonaConvGrads(ctypes.Structure):
    _fields_ = [(n, _c_void_p) for n in ("dk3", "db3", "dk5", "db5", "dk7", "db7", "dP", "dbp", "db1", "dfreq", "dne_w1", "dne_b1", "dne_w2", "dne_b2")]


class MonaConvDesc(ctypes.Structure):
    _fields_ = [("h", _c_void_p), ("g", _c_void_p), ("dg", _c_void_p), ("dh", _c_void_p),
                ("w", MonaConvWeights), ("gr", MonaConvGrads),
                ("B", _c_int), ("N", _c_int), ("H", _c_int), ("W", _c_int), ("C", _c_int), ("has_cls", _c_int),
                ("drop_p", _c_float), ("seed", _u64), ("dtype", _c_int), ("force_simt", _c_int)]


class MonaParams(ctypes.Structure):   # ngu_mona_params
    _fields_ = [(n, _c_void_p) for n in ("w1", "b1", "w2", "b2", "ln_w", "ln_b", "gamma", "gammax")] + [("conv", MonaConvWeights)]


class MonaDerived(ctypes.Structure):  # ngu_mona_derived
    _fields_ = [(n, _c_void_p) for n in ("wab", "wcat_t", "w2", "w2_t", "ca", "cb", "kc", "bc", "pb", "bp")]


class MonaPrepItem(ctypes.Structure):
    _fields_ = [("p", MonaParams), ("d", MonaDerived)]


class MonaStageDesc(ctypes.Structure):
    _fields_ = [("d", MonaDerived), ("x", _c_void_p), ("h", _c_void_p), ("hA", _c_void_p), ("g", _c_void_p),
                ("mean", _c_void_p), ("rstd", _c_void_p), ("dg", _c_void_p), ("dhcat", _c_void_p), ("rowab", _c_void_p),
                ("ws", _c_void_p), ("dP", _c_void_p), ("dbp", _c_void_p),
                ("B", _c_int), ("N", _c_int), ("H", _c_int), ("W", _c_int), ("D", _c_int), ("has_cls", _c_int),
                ("eps", _c_float), ("drop_p", _c_float), ("seed", _u64)]


class MonaGrads(ctypes.Structure):    # ngu_mona_grads
    _fields_ = [(n, _c_void_p) for n in ("dw1", "db1", "dln_w", "dln_b", "dgamma", "dgammax",
                                         "dk3", "db3", "dk5", "db5", "dk7", "db7", "dfreq")]


class AttnDesc(ctypes.Structure):
    _fields_ = [("q", _c_void_p), ("q_bs", _i64), ("q_ts", _i64),
                ("k", _c_void_p), ("k_bs", _i64), ("k_ts", _i64),
                ("v", _c_void_p), ("v_bs", _i64), ("v_ts", _i64),
                ("o", _c_void_p), ("o_bs", _i64), ("o_ts", _i64),
                ("lse", _c_void_p), ("d_o", _c_void_p),
                ("dq", _c_void_p), ("dk", _c_void_p), ("dv", _c_void_p),
                ("B", _c_int), ("H", _c_int), ("N", _c_int), ("S", _c_int), ("dh", _c_int),
                ("scale", _c_float), ("causal", _c_int), ("dtype", _c_int), ("impl", _c_int), ("kv_len", _c_void_p), ("ws", _c_void_p)]


class InfoNceDesc(ctypes.Structure):
    _fields_ = [("ihat", _c_void_p), ("that", _c_void_p), ("dihat", _c_void_p), ("dthat", _c_void_p),
                ("loss", _c_void_p), ("ws", _c_void_p),
                ("Bg", _c_int), ("Bl", _c_int), ("r0", _c_int), ("E", _c_int), ("temperature", _c_float),
                ("ihat16", _c_void_p), ("that16", _c_void_p), ("ihat16_t", _c_void_p), ("that16_t", _c_void_p), ("g_ws", _c_void_p)]


class AdamWDesc(ctypes.Structure):
    _fields_ = [("param", _c_void_p), ("grad", _c_void_p), ("m", _c_void_p), ("v", _c_void_p), ("n", _i64),
                ("lr", _c_float), ("beta1", _c_float), ("beta2", _c_float), ("eps", _c_float), ("weight_decay", _c_float),
                ("step", _c_int), ("max_norm", _c_float), ("gsq", _c_void_p), ("loss", _c_void_p), ("zero_grad", _c_int),
                ("state", _c_void_p), ("lr_min", _c_float), ("t_max", _c_int)]


# every symbol include/ngu_b200.h declares: name -> (restype, argtypes)
def _P(t):
    return ctypes.POINTER(t)


PROTOTYPES = {
    "ngu_version": (_c_int, []),
    "ngu_last_error": (ctypes.c_char_p, []),
    "ngu_launch_count": (_i64, []),
    "ngu_selftest_device": (_c_int, []),
    "ngu_gemm": (_c_int, [_P(GemmDesc), _c_void_p]),
    "ngu_ln_fwd": (_c_int, [_P(LnDesc), _c_void_p]),
    "ngu_ln_bwd": (_c_int, [_P(LnBwdDesc), _c_void_p]),
    "ngu_mona_pre_bwd": (_c_int, [_P(MonaPreBwdDesc), _c_void_p]),
    "ngu_mona_conv_fwd": (_c_int, [_P(MonaConvDesc), _c_void_p]),
    "ngu_mona_conv_bwd": (_c_int, [_P(MonaConvDesc), _c_void_p]),
    "ngu_mona_ws_floats": (_i64, [_c_int]),
    "ngu_mona_prep": (_c_int, [_c_void_p, _c_int, _c_int, _c_void_p]),
    "ngu_mona_fwd_stage": (_c_int, [_P(MonaStageDesc), _c_void_p]),
    "ngu_mona_bwd_stage": (_c_int, [_P(MonaStageDesc), _c_void_p]),
    "ngu_mona_finish": (_c_int, [_P(MonaParams), _P(MonaGrads), _c_void_p, _c_int, _c_void_p]),
    "ngu_attn_fwd": (_c_int, [_P(AttnDesc), _c_void_p]),
    "ngu_attn_bwd": (_c_int, [_P(AttnDesc), _c_void_p]),
    "ngu_infonce_normalize": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p]),
    "ngu_infonce_core": (_c_int, [_P(InfoNceDesc), _c_void_p]),
    "ngu_infonce_normalize_bwd": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p]),
    "ngu_wgrad": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p]),
    "ngu_colsum": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_int, _c_int, _c_int, _c_void_p]),
    "ngu_dropout": (_c_int, [_c_void_p, _c_void_p, _i64, _c_float, _u64, _c_int, _c_int, _c_void_p]),
    "ngu_sqnorm": (_c_int, [_c_void_p, _i64, _c_void_p, _c_void_p]),
    "ngu_adamw_step": (_c_int, [_P(AdamWDesc), _c_void_p]),
    "ngu_guard_tick": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_void_p]),
    "ngu_set_seed_counter": (_c_int, [_c_void_p]),
    "ngu_zero_shot_prototypes": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p]),
    "ngu_zero_shot_score": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_float, _c_int, _c_void_p]),
    "ngu_kv_len": (_c_int, [_c_void_p, _i64, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p]),
    "ngu_patchify": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p]),
    "ngu_assemble_tokens": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p]),
    "ngu_embed_tokens": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p]),
    "ngu_cast_f32": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_float, _c_int, _c_void_p]),
    "ngu_cast_f32_batch": (_c_int, [_c_void_p, _c_int, _c_int, _c_void_p]),
}


class NguError(RuntimeError):
    pass


_lib = None


def lib():
    """Load (once) and return the ctypes handle; raises if the extension was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NguError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback.")
        h = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(h, name)  # AttributeError here = header/library mismatch: fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().ngu_last_error().decode("utf-8", "replace")
        raise NguError(f"{what} failed (code {rc}): {msg}")


def launch_count():
    return int(lib().ngu_launch_count())
