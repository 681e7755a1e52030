"""Segmented CUDA-graph capture of a step that contains NCCL collectives.

Capturing NCCL calls into a CUDA graph hangs on this stack (torch 2.11 / NCCL 2.28, both capture-error modes; DESIGN.md
section 5a), so a data-parallel step is recorded as graph SEGMENTS with the collectives launched eagerly between them:

    [graph: towers forward, feature normalise] -> all_gather (NCCL) -> [graph: loss, backward] -> all_reduce (NCCL)
    -> [graph: clip + AdamW + guard]

`collective(fn)` is the one hook the library code needs: outside a capture it just runs `fn`; inside one it ends the current
segment, runs and remembers `fn`, and opens the next segment on the same memory pool (so tensors made in one segment stay
valid in the next).  A replay walks the list: graph launches and NCCL calls alternate on the caller's stream.
"""
import torch

ACTIVE = None


def collective(fn):
    if ACTIVE is None:
        fn()
    else:
        ACTIVE.cut(fn)


class SegmentedCapture:
    def __init__(self, error_mode="thread_local"):
        self.items = []          # CUDAGraph objects and callables, in replay order
        self.pool = None
        self.g = None
        self.error_mode = error_mode

    def _begin(self):
        self.g = torch.cuda.CUDAGraph()
        if self.pool is None:
            self.pool = torch.cuda.graph_pool_handle()   # one private pool for every segment
        self.g.capture_begin(pool=self.pool, capture_error_mode=self.error_mode)

    def _end(self):
        self.g.capture_end()
        self.items.append(self.g)
        self.g = None

    def __enter__(self):
        global ACTIVE
        if ACTIVE is not None:
            raise RuntimeError("nested segmented capture")
        ACTIVE = self
        self._begin()
        return self

    def __exit__(self, exc_type, exc, tb):
        global ACTIVE
        ACTIVE = None
        if self.g is not None:
            if exc_type is None:
                self._end()
            else:
                try:
                    self.g.capture_end()
                except Exception:
                    pass
        return False

    def cut(self, fn):
        self._end()
        fn()                     # executed now too (on not-yet-computed buffers): every rank must issue the same collectives
        self.items.append(fn)
        self._begin()

    def replay(self):
        for it in self.items:
            if isinstance(it, torch.cuda.CUDAGraph):
                it.replay()
            else:
                it()

    @property
    def segments(self):
        return sum(isinstance(it, torch.cuda.CUDAGraph) for it in self.items)
