"""Drop-in at model level, checked on the reference's REAL class (modeling_bitllama.py:25,229-231,451-454,1512):
`replace_bitlinear` on a `BitLlamaForCausalLMInf` must swap all 7 projections per layer for `BitLinearB200`, keep the
very same Parameter objects and leave `state_dict()` (keys, order, tensors) untouched.

Runs only where the reference tree is mounted (the build container); the GPU box has no /root/reference, and the
numerical side of the swap is covered there by tests/test_decoder_gpu.py against fixtures the reference produced."""
import os
import sys
import types
from pathlib import Path

import pytest
import torch

from onebit_b200 import BitLinearB200, replace_bitlinear

REF = Path(os.environ.get("ONEBIT_REFERENCE", "/root/reference"))
pytestmark = pytest.mark.skipif(not (REF / "transformers/src/transformers/models/bitllama").is_dir(),
                                reason="reference tree not mounted")

TINY = dict(vocab_size=384, hidden_size=256, intermediate_size=768, num_hidden_layers=2, num_attention_heads=2,
            num_key_value_heads=2, max_position_embeddings=256, rms_norm_eps=1e-6, hidden_act="silu",
            rope_theta=10000.0, pad_token_id=0, bos_token_id=1, eos_token_id=None, tie_word_embeddings=False)


@pytest.fixture(scope="module")
def ref_classes():
    saved_path, saved_mods = list(sys.path), {k: v for k, v in sys.modules.items() if k == "transformers" or k.startswith("transformers.")}
    for k in saved_mods:
        del sys.modules[k]
    stub = types.ModuleType("transformers.dependency_versions_check")  # only enforces a tokenizers version pin
    stub.dep_version_check = lambda *a, **k: None
    sys.modules["transformers.dependency_versions_check"] = stub
    sys.path.insert(0, str(REF / "transformers/src"))
    try:
        from transformers import BitLlamaConfig, BitLlamaForCausalLMInf
        from transformers.models.bitnet import BitLinearInf
        yield BitLlamaConfig, BitLlamaForCausalLMInf, BitLinearInf
    finally:
        for k in [k for k in sys.modules if k == "transformers" or k.startswith("transformers.")]:
            del sys.modules[k]
        sys.modules.update(saved_mods)
        sys.path[:] = saved_path


def test_replace_bitlinear_on_the_real_reference_model(ref_classes):
    BitLlamaConfig, BitLlamaForCausalLMInf, BitLinearInf = ref_classes
    torch.manual_seed(0)
    model = BitLlamaForCausalLMInf(BitLlamaConfig(**TINY)).eval()
    before = model.state_dict()
    keys_before = list(before.keys())
    params_before = {n: p for n, p in model.named_parameters()}
    n_ref = sum(isinstance(m, BitLinearInf) for m in model.modules())
    assert n_ref == 7 * TINY["num_hidden_layers"]

    ln_before = model.model.layers[0].self_attn.q_proj.layernorm
    assert replace_bitlinear(model) == n_ref
    assert not any(isinstance(m, BitLinearInf) for m in model.modules())
    assert model.model.layers[0].self_attn.q_proj.layernorm is ln_before  # the reference's nn.LayerNorm attribute is adopted
    for layer in model.model.layers:
        for mod in (layer.self_attn.q_proj, layer.self_attn.k_proj, layer.self_attn.v_proj, layer.self_attn.o_proj,
                    layer.mlp.gate_proj, layer.mlp.up_proj, layer.mlp.down_proj):
            assert isinstance(mod, BitLinearB200)
    # lm_head stays a dense nn.Linear (modeling_bitllama.py:1519)
    assert type(model.lm_head) is torch.nn.Linear

    after = model.state_dict()
    assert list(after.keys()) == keys_before
    for k in keys_before:
        assert after[k].data_ptr() == before[k].data_ptr() and after[k].dtype == before[k].dtype
    params_after = {n: p for n, p in model.named_parameters()}
    assert params_after.keys() == params_before.keys()
    for n, p in params_after.items():
        assert p is params_before[n], n
    assert replace_bitlinear(model) == 0  # idempotent

    # a reference checkpoint loads into the swapped model unchanged (same keys / shapes / dtypes)
    ref2 = BitLlamaForCausalLMInf(BitLlamaConfig(**TINY)).eval()
    missing, unexpected = model.load_state_dict(ref2.state_dict(), strict=True)
    assert not missing and not unexpected
    # int8 sign bytes survive .half() like in the reference's loading path
    model.half()
    assert model.model.layers[0].self_attn.q_proj.weight.dtype == torch.int8
    assert model.model.layers[0].self_attn.q_proj.weight_scale.dtype == torch.float16


def test_swapped_model_refuses_cpu_forward_loudly(ref_classes):
    BitLlamaConfig, BitLlamaForCausalLMInf, _ = ref_classes
    model = BitLlamaForCausalLMInf(BitLlamaConfig(**TINY)).float().eval()
    replace_bitlinear(model)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(torch.tensor([[1, 2, 3]]))
