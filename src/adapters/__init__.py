"""Drop-in replacement for the reference's src/adapters/__init__.py.

The reference file imports names that do not exist in its own tree (`FractionalMona`, `.prompt_tuning`;
reference src/adapters/__init__.py:21-39) and therefore fails to import as shipped.  This
version exports every adapter symbol that HAS an implementation in the reference and that is built here."""
from nextgen_uia_b200.adapters.mona import (  # noqa: F401
    BaselineMona,
    BaselineMonaOp,
    NoiseAwareMona,
    NoiseAwareMonaOp,
    FreqEnhancedMona,
    FreqEnhancedMonaOp,
    HybridNoiseFreqMona,
    HybridNoiseFreqMonaOp,
    BatchFirstMonaWrapper,
    inject_mona_variant_to_clip,
    inject_mona_variant_to_open_clip,
)
from nextgen_uia_b200.adapters.lora import (  # noqa: F401
    LoRALayer,
    LinearLoRA,
    PlainMultiheadAttentionLoRA,
    inject_lora_to_clip,
    inject_lora_to_biomedclip,
)

__all__ = [
    "BaselineMona", "BaselineMonaOp", "NoiseAwareMona", "NoiseAwareMonaOp", "FreqEnhancedMona", "FreqEnhancedMonaOp",
    "HybridNoiseFreqMona", "HybridNoiseFreqMonaOp", "BatchFirstMonaWrapper",
    "inject_mona_variant_to_clip", "inject_mona_variant_to_open_clip",
    "LoRALayer", "LinearLoRA", "PlainMultiheadAttentionLoRA",
    "inject_lora_to_clip", "inject_lora_to_biomedclip",
]
