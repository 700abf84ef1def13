"""CPU checks of the host-side mirror of BitLlamaForCausalLMInf: RoPE tables bit-identical to the reference's
LlamaRotaryEmbedding (golden), synthetic state dict uses the reference's key names / dtypes / shapes."""
import numpy as np
import torch

from onebit_b200.bitllama import rope_tables, synthetic_state_dict


def test_rope_tables_bit_identical_to_reference(golden_dir):
    z = np.load(golden_dir / "rope_tables.npz")
    cos, sin = rope_tables(128, 512, 10000.0)
    # the reference concatenates two identical halves (modeling_bitllama.py:106-108); we keep the first
    np.testing.assert_array_equal(cos.numpy(), z["cos"][:, :64])
    np.testing.assert_array_equal(sin.numpy(), z["sin"][:, :64])
    np.testing.assert_array_equal(z["cos"][:, :64], z["cos"][:, 64:])


def test_synthetic_state_dict_uses_reference_keys(golden_dir):
    z = np.load(golden_dir / "tiny_model.npz")
    cfg = {k: v for k, v in zip(z["config_keys"], z["config_vals"])}
    config = {k: int(cfg[k]) for k in ("hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads",
                                       "vocab_size")}
    ref = {k[4:]: z[k] for k in z.files if k.startswith("sd::")}
    sd = synthetic_state_dict(config, seed=3)
    assert set(sd) == set(ref)  # identical key set to BitLlamaForCausalLMInf.state_dict()
    for k, v in sd.items():
        assert tuple(v.shape) == ref[k].shape, k
        if k.endswith("_proj.weight"):
            assert v.dtype == torch.int8 and ref[k].dtype == np.int8
