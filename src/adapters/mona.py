from nextgen_uia_b200.adapters.mona import *  # noqa: F401,F403
from nextgen_uia_b200.adapters.mona import (BaselineMona, BaselineMonaOp, NoiseAwareMona, NoiseAwareMonaOp,  # noqa: F401
                                            FreqEnhancedMona, FreqEnhancedMonaOp, HybridNoiseFreqMona, HybridNoiseFreqMonaOp,
                                            BatchFirstMonaWrapper,
                                            inject_mona_variant_to_clip, inject_mona_variant_to_open_clip)
