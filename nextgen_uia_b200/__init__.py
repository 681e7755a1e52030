"""nextgen_uia_b200 — B200-native (sm_100a) implementation of NextGen-UIA's adapter fine-tuning hot path.

Layout (only what the hot path needs):
  csrc/                 hand-written CUDA kernels + the C ABI (include/ngu_b200.h) -> libngu_b200.so
  _lib.py, ops.py       ctypes binding and tensor-level wrappers (no torch math, no CPU fallback)
  adapters/{mona,lora}  drop-in mirrors of the reference's src/adapters modules
  losses.py             drop-in mirror of src/losses (InfoNCELoss), with the data-parallel gather
  vit.py, biomedclip.py timm/open_clip-shaped towers whose blocks run on the kernels
  dp.py                 one-process-per-GPU data-parallel training step (NCCL)
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"
