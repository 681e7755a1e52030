from nextgen_uia_b200.adapters.mona import *  # noqa: F401,F403
from nextgen_uia_b200.adapters.mona import (BaselineMona, BaselineMonaOp, BatchFirstMonaWrapper,  # noqa: F401
                                            inject_mona_variant_to_clip, inject_mona_variant_to_open_clip)
