#!/usr/bin/env python
"""BASELINE.json configs[2]: LLaMA-7B OneBit, prefill 2048 + decode 128, batch 8, one B200.
    python tools/bench_config3.py [--batch 8] [--prompt 2048] [--new 128] [--model 7b]
Prints one JSON line: prompt-pass time / tok/s / achieved BitLinear TFLOP/s, decode ms per step / tok/s, end-to-end."""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from onebit_b200 import LLAMA2_13B, LLAMA_7B, BitLlamaDecoderB200, synthetic_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--prompt", type=int, default=2048)
ap.add_argument("--new", type=int, default=128)
ap.add_argument("--model", default="7b")
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
cfg = LLAMA_7B if args.model == "7b" else LLAMA2_13B
B, T, NEW = args.batch, args.prompt, args.new
sd = synthetic_state_dict(cfg, seed=0)
dec = BitLlamaDecoderB200(cfg, sd, max_seq_len=T + NEW + 16, max_batch=B)
del sd
prompt = torch.randint(3, cfg["vocab_size"], (B, T), generator=torch.Generator().manual_seed(1))
dec.prefill(prompt)  # warm-up (allocates the workspace)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
pre_ms = []
for _ in range(args.reps):
    e[0].record()
    dec.prefill(prompt)
    e[1].record()
    torch.cuda.synchronize()
    pre_ms.append(e[0].elapsed_time(e[1]))
for _ in range(3):
    dec.step()
torch.cuda.synchronize()
dec.prefill(prompt)
e[2].record()
for _ in range(NEW):
    dec.step()
e[3].record()
torch.cuda.synchronize()
dec_ms = e[2].elapsed_time(e[3])
H, I, L = cfg["hidden_size"], cfg["intermediate_size"], cfg["num_hidden_layers"]
flops = 2.0 * B * T * L * (4 * H * H + 3 * H * I)
attn_flops = 4.0 * B * L * cfg["num_attention_heads"] * 128 * T * (T + 1) / 2
best = min(pre_ms)
peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
tpeak = peaks.get("bf16_tflops_sustained", 1400.0)
print(json.dumps({"config": f"{args.model} prefill {T} + decode {NEW}, batch {B}", "prefill_ms": best, "prefill_ms_all": pre_ms,
                  "prefill_tok_s": B * T / (best * 1e-3), "bitlinear_tflops_over_whole_prompt_pass": flops / (best * 1e-3) / 1e12,
                  "tensor_frac_of_sustained_peak": flops / (best * 1e-3) / 1e12 / tpeak, "attention_tflops_if_alone": attn_flops / 1e12,
                  "decode_ms_per_step": dec_ms / NEW, "decode_tok_s": B * NEW / (dec_ms * 1e-3), "decode_launches_per_step": dec.launches_per_step(),
                  "end_to_end_s": (best + dec_ms) * 1e-3, "status": dec.status()}), flush=True)
