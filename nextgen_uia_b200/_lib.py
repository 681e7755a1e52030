"""ctypes binding of libngu_b200.so (C ABI in include/ngu_b200.h).

The library is the product: there is no CPU or PyTorch fallback.  If the shared object is missing
or no sm_100 device is usable, calls raise instead of silently computing something else.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libngu_b200.so")

NGU_BF16, NGU_F32 = 0, 1
ACT_NONE, ACT_GELU, ACT_QUICKGELU = 0, 1, 2
AUX_NONE, AUX_RESIDUAL, AUX_DACT = 0, 1, 2

_c_void_p, _c_int, _c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_float


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
    ]


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
        h.ngu_version.restype = _c_int
        h.ngu_last_error.restype = ctypes.c_char_p
        h.ngu_launch_count.restype = ctypes.c_int64
        h.ngu_selftest_device.restype = _c_int
        h.ngu_gemm.argtypes = [ctypes.POINTER(GemmDesc), _c_void_p]
        h.ngu_gemm.restype = _c_int
        _lib = h
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().ngu_last_error().decode("utf-8", "replace")
        raise NguError(f"{what} failed (code {rc}): {msg}")


def launch_count():
    return int(lib().ngu_launch_count())
