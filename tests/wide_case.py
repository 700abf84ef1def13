"""Helper run as a subprocess by tests/test_decoder_gpu.py (the decode-path switches ONEBIT_PERSIST / ONEBIT_FUSED are
read once per process): decoder logits at real LLaMA widths against the pinned CPU port of the reference.

    python tests/wide_case.py <7b|13b> <layers> <tokens> <batch> <f16|f32>

Prints one JSON line: rel-L2 of the logits vs oracle/ref_port.py (fp32, the reference's op sequence
modeling_bitllama.py:856-930,1512-1611), arg-max agreement, launches per step, decoder status."""
import json
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from onebit_b200 import LLAMA2_13B, LLAMA_7B, BitLlamaDecoderB200, synthetic_state_dict  # noqa: E402
from oracle import oracle, ref_port  # noqa: E402


def main(model, layers, tokens, batch, pdt):
    config = dict(LLAMA_7B if model == "7b" else LLAMA2_13B, num_hidden_layers=layers)
    sd = synthetic_state_dict(config, seed=3, param_dtype=torch.float32)
    if os.environ.get("ONEBIT_WIDE_OUTLIERS") == "1":  # a few massive-activation channels in the residual stream
        emb = sd["model.embed_tokens.weight"]
        emb[:, [7, 1000, 2049]] *= 60.0
    ids = torch.randint(3, config["vocab_size"], (batch, tokens), generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        want, _ = ref_port.RefPortModel(config, sd).forward(ids)
    dec = BitLlamaDecoderB200(config, sd, max_seq_len=max(64, tokens), max_batch=batch,
                              param_dtype=torch.float32 if pdt == "f32" else torch.float16)
    if os.environ.get("ONEBIT_WIDE_PREFILL") == "1":
        got = dec.prefill(ids, all_logits=True).cpu().numpy()
    else:
        got = dec.forward_tokens(ids).cpu().numpy()
    out = {"rel_l2": float(oracle.rel_l2(got, want.numpy())),
           "argmax_agree": float((got.argmax(-1) == want.numpy().argmax(-1)).mean()),
           "launches": dec.launches_per_step(), "persistent": bool(dec.persistent), "status": dec.status()}
    dec.close()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5])
