from nextgen_uia_b200.adapters.lora import (LoRALayer, LinearLoRA, PlainMultiheadAttentionLoRA,  # noqa: F401
                                            inject_lora_to_clip, inject_lora_to_biomedclip)
