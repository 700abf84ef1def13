"""onebit_b200 — B200-native (sm_100a) drop-in for OneBit's 1-bit linear layer.

Host-side mirror of the reference interface (xuyuzhuang11/OneBit,
transformers/src/transformers/models/bitnet.py:71-122) over the C ABI in include/onebit_b200.h.
"""
from .bitlinear import (BitLinearB200, bitlinear_forward, bitlinear_matvec, pack_signs, replace_bitlinear,
                        scale_layernorm, unpack_signs)

from .bitllama import BitLlamaDecoderB200, synthetic_state_dict, LLAMA_7B, LLAMA2_13B  # noqa: E402

__all__ = ["BitLlamaDecoderB200", "synthetic_state_dict", "LLAMA_7B", "LLAMA2_13B", "BitLinearB200", "bitlinear_forward", "bitlinear_matvec", "scale_layernorm", "pack_signs",
           "unpack_signs", "replace_bitlinear"]
__version__ = "0.1.0"
