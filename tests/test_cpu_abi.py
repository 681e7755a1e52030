"""CPU-only checks: the C-ABI library loads and exports every symbol include/ngu_b200.h declares, the
product path refuses to run without a GPU (no CPU fallback), error codes surface as exceptions."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "ngu_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ngu_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from nextgen_uia_b200 import _lib as L
    h = ctypes.CDLL(L.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(h, s), f"{s} declared in include/ngu_b200.h but not exported"
    assert set(L.PROTOTYPES) == set(syms), set(L.PROTOTYPES) ^ set(syms)
    assert L.lib().ngu_version() == 100


def test_struct_sizes_match_header():
    """ctypes mirrors must have the C layout (compile a probe with gcc against the header)."""
    import subprocess, tempfile
    from nextgen_uia_b200 import _lib as L
    names = {"ngu_gemm_desc": L.GemmDesc, "ngu_ln_desc": L.LnDesc, "ngu_ln_bwd_desc": L.LnBwdDesc,
             "ngu_mona_pre_bwd_desc": L.MonaPreBwdDesc, "ngu_mona_conv_desc": L.MonaConvDesc, "ngu_attn_desc": L.AttnDesc,
             "ngu_infonce_desc": L.InfoNceDesc, "ngu_adamw_desc": L.AdamWDesc, "ngu_cast_item": L.CastItem,
             "ngu_mona_params": L.MonaParams, "ngu_mona_derived": L.MonaDerived, "ngu_mona_prep_item": L.MonaPrepItem,
             "ngu_mona_stage_desc": L.MonaStageDesc, "ngu_mona_grads": L.MonaGrads}
    prog = '#include <stdio.h>\n#include "ngu_b200.h"\nint main(){' + "".join(
        f'printf("{n} %zu\\n", sizeof({n}));' for n in names) + "return 0;}"
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "p.c")
        open(c, "w").write(prog)
        exe = os.path.join(td, "p")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        out = subprocess.check_output([exe], text=True)
    for line in out.strip().splitlines():
        n, sz = line.split()
        assert ctypes.sizeof(names[n]) == int(sz), (n, ctypes.sizeof(names[n]), sz)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from nextgen_uia_b200 import _lib as L, ops
    assert L.lib().ngu_selftest_device() != 0
    with pytest.raises(L.NguError):
        ops.gemm(torch.zeros(8, 8), torch.zeros(8, 8))
    from src.adapters import BaselineMona
    m = BaselineMona(768, 64)
    with pytest.raises(L.NguError):
        m(torch.zeros(197, 1, 768), (14, 14))
    from src.losses import InfoNCELoss
    with pytest.raises(L.NguError):
        InfoNCELoss()(torch.zeros(4, 512), torch.zeros(4, 512))


def test_null_descriptor_is_an_error_code():
    from nextgen_uia_b200 import _lib as L
    rc = L.lib().ngu_gemm(None, None)
    assert rc == -5
    assert b"null" in L.lib().ngu_last_error()
