"""flasht5_b200 -- the hot path of catie-aq/flashT5 (FlashAttention-2 with additive T5 bias, plus
the RMSNorm and cross-entropy ops) as hand-written sm_100a CUDA behind a C ABI, with the
reference's own Python operator surface on top.

    from flasht5_b200 import flash_attention_v2_bias, fast_rms_layernorm, cross_entropy_loss
    from flasht5_b200 import RelativePositionalEncoding, flash_attention_v2_rpe   # bias producer / in-kernel bias

The layout follows the reference's src/model/ops/: one module per operator, same function and
class names, same argument meaning.  No Triton, no multi-backend dispatch, no CPU fallback.
"""
from .flash_attention_v2_bias import FlashAttentionAdditiveBias, SharedBiasGrad, flash_attention_v2_bias, flash_attention_v2_bias_shared
from .rms_norm import Fast_RMS_Layernorm, fast_rms_layernorm
from .cross_entropy_loss import CrossEntropyLoss, cross_entropy_loss
from .positional_encoding import RelativePositionalEncoding
from .flash_attention_rpe import FlashAttentionRPE, flash_attention_v2_rpe
from .adamw_scaled import AdamWScale

__all__ = [
    "flash_attention_v2_bias", "FlashAttentionAdditiveBias", "flash_attention_v2_bias_shared", "SharedBiasGrad",
    "fast_rms_layernorm", "Fast_RMS_Layernorm",
    "cross_entropy_loss", "CrossEntropyLoss",
    "RelativePositionalEncoding",
    "flash_attention_v2_rpe", "FlashAttentionRPE",
    "AdamWScale",
]
__version__ = "0.1.0"
