"""Pins oracle/ref_port.py (the op-for-op CPU port used as the timed reference arm) against outputs of the
reference's own BitLlamaForCausalLMInf (tests/golden/tiny_model.npz)."""
import numpy as np
import torch

from oracle import oracle, ref_port


def _load(golden_dir):
    z = np.load(golden_dir / "tiny_model.npz")
    cfg = {k: v for k, v in zip(z["config_keys"], z["config_vals"])}
    config = {k: (float(cfg[k]) if k in ("rms_norm_eps", "rope_theta") else int(cfg[k]))
              for k in ("hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads", "vocab_size",
                        "rms_norm_eps", "rope_theta")}
    sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd::")}
    return config, sd, z


def test_port_full_sequence_logits(golden_dir):
    config, sd, z = _load(golden_dir)
    model = ref_port.RefPortModel(config, sd)
    with torch.no_grad():
        logits, _ = model.forward(torch.from_numpy(z["input_ids"]))
    assert oracle.rel_l2(logits.numpy(), z["logits"]) < 1e-5


def test_port_decode_with_cache_equals_full_sequence(golden_dir):
    config, sd, z = _load(golden_dir)
    model = ref_port.RefPortModel(config, sd)
    ids = torch.from_numpy(z["input_ids"])[:, :12]
    with torch.no_grad():
        full, _ = model.forward(ids)
        past, outs = None, []
        for i in range(ids.shape[1]):
            lg, past = model.forward(ids[:, i:i + 1], past, pos=i)
            outs.append(lg)
    step = torch.cat(outs, dim=1)
    assert oracle.rel_l2(step.numpy(), full.numpy()) < 1e-5


def test_port_bitlinear_matches_c_oracle():
    case = oracle.synth_case(3, 256, 256, 8, with_bias=True)
    got = ref_port.bitlinear_forward(torch.from_numpy(case["x"]), torch.from_numpy(case["packed"]),
                                     torch.from_numpy(case["g"]), torch.from_numpy(case["h"]),
                                     torch.from_numpy(case["bias"])).numpy()
    want = oracle.bitlinear_forward_c(case["x"], case["packed"], case["g"], case["h"], case["bias"])
    assert oracle.rel_l2(got, want) < 2e-6


def test_port_reproduces_the_512_token_perplexity_fixture(golden_dir):
    """tests/golden/tiny_ppl512.npz: evaluation/lm_eval.py:99-124 on one fixed 512-token window, produced by the
    reference's own BitLlamaForCausalLMInf. The port must land on the same number to 3 decimals."""
    config, sd, _ = _load(golden_dir)
    z = np.load(golden_dir / "tiny_ppl512.npz")
    ids = torch.from_numpy(z["input_ids"])
    assert ids.shape == (1, 512)
    model = ref_port.RefPortModel(config, sd)
    with torch.no_grad():
        logits, _ = model.forward(ids)
    loss = torch.nn.functional.cross_entropy(logits[0, :-1].double(), ids[0, 1:])
    ppl = float(torch.exp(loss))
    assert abs(ppl - float(z["ppl64"])) < 5e-4
    assert abs(ppl - float(z["ppl"])) < 5e-4
    assert oracle.rel_l2(logits[0, -1].numpy(), z["last_logits"]) < 1e-5
