from .mona import (BaselineMona, BaselineMonaOp, NoiseAwareMona, NoiseAwareMonaOp, FreqEnhancedMona, FreqEnhancedMonaOp,  # noqa: F401
                   HybridNoiseFreqMona, HybridNoiseFreqMonaOp, BatchFirstMonaWrapper, inject_mona_variant_to_clip,
                   inject_mona_variant_to_open_clip)
from .lora import (LoRALayer, LinearLoRA, PlainMultiheadAttentionLoRA, inject_lora_to_clip,  # noqa: F401
                   inject_lora_to_biomedclip)
