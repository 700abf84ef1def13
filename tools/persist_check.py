#!/usr/bin/env python
"""Bring-up / regression tool for the persistent decode step (run on the GPU box):
    python tools/persist_check.py tiny|wide7b|wide13b|time7b|time13b [...]
Each section prints one JSON line. Sections are independent; run each under `timeout`."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from onebit_b200 import LLAMA2_13B, LLAMA_7B, BitLlamaDecoderB200, synthetic_state_dict  # noqa: E402
from oracle import oracle, ref_port  # noqa: E402


def load_tiny():
    z = np.load(ROOT / "tests" / "golden" / "tiny_model.npz")
    cfg = {k: v for k, v in zip(z["config_keys"], z["config_vals"])}
    config = {k: (float(cfg[k]) if k in ("rms_norm_eps", "rope_theta") else int(cfg[k]))
              for k in ("hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads", "vocab_size",
                        "rms_norm_eps", "rope_theta")}
    sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd::")}
    return config, sd, z


def sec_tiny():
    config, sd, z = load_tiny()
    out = {"section": "tiny"}
    for pd, nm in ((torch.float32, "f32"), (torch.float16, "f16")):
        dec = BitLlamaDecoderB200(config, sd, max_seq_len=256, max_batch=2, param_dtype=pd, use_graph=(pd == torch.float16))
        ids = torch.from_numpy(z["input_ids"])
        logits = dec.forward_tokens(ids).cpu().numpy()
        out[f"persistent_{nm}"] = dec.persistent
        out[f"status_{nm}"] = dec.status()
        out[f"logits_rel_l2_{nm}"] = float(oracle.rel_l2(logits, z["logits"]))
        out[f"finite_{nm}"] = bool(np.isfinite(logits).all())
        ppl = dec.perplexity(ids)
        out[f"ppl_{nm}"] = ppl
        out["ppl_ref"] = float(z["ppl"])
        prompt = torch.from_numpy(z["prompt"])
        want = z["generated"]
        got = dec.generate(prompt, max_new_tokens=want.shape[1] - prompt.shape[1]).cpu().numpy()
        out[f"greedy_agree_{nm}"] = float((got == want).mean())
        # batch 1 must equal row 0 of batch 2 bit for bit
        d1 = BitLlamaDecoderB200(config, sd, max_seq_len=256, max_batch=1, param_dtype=pd, use_graph=False)
        a = d1.forward_tokens(ids[:1, :16])
        b = dec.forward_tokens(ids[:, :16])
        out[f"batch_invariant_{nm}"] = bool(torch.equal(a[0], b[0]))
        d1.close()
        dec.close()
    print(json.dumps(out), flush=True)


def wide(config, name, layers, tokens, batch=1):
    config = dict(config, num_hidden_layers=layers)
    sd = synthetic_state_dict(config, seed=3, param_dtype=torch.float32)
    gen = torch.Generator().manual_seed(5)
    ids = torch.randint(3, config["vocab_size"], (batch, tokens), generator=gen)
    model = ref_port.RefPortModel(config, sd)
    t0 = time.time()
    with torch.no_grad():
        want, _ = model.forward(ids)
    t_ref = time.time() - t0
    out = {"section": name, "layers": layers, "tokens": tokens, "batch": batch, "ref_s": t_ref}
    for pd, nm in ((torch.float32, "f32"), (torch.float16, "f16")):
        dec = BitLlamaDecoderB200(config, sd, max_seq_len=64, max_batch=batch, param_dtype=pd)
        got = dec.forward_tokens(ids).cpu().numpy()
        if dec.status() != 0:
            tr = dec.read_trace()
            out[f"abort_info_{nm}"] = {"cta": int(tr[0, 0, 8]), "tid": int(tr[0, 0, 9]), "site_layer": int(tr[0, 0, 10]) // 16,
                                      "site_stage": int(tr[0, 0, 10]) % 16}
        out[f"persistent_{nm}"] = dec.persistent
        out[f"status_{nm}"] = dec.status()
        out[f"logits_rel_l2_{nm}"] = float(oracle.rel_l2(got, want.numpy()))
        out[f"argmax_agree_{nm}"] = float((got.argmax(-1) == want.numpy().argmax(-1)).mean())
        dec.close()
    print(json.dumps(out), flush=True)


def timing(config, name, batch=1, steps=64, prompt_len=16):
    sd = synthetic_state_dict(config, seed=0)
    dec = BitLlamaDecoderB200(config, sd, max_seq_len=prompt_len + steps * 3 + 64, max_batch=batch)
    del sd
    gen = torch.Generator().manual_seed(1234)
    prompt = torch.randint(3, config["vocab_size"], (batch, prompt_len), generator=gen)
    dec.reset(prompt[:, 0])
    ids = prompt.cuda()
    for i in range(prompt_len):
        dec.step(ids[:, i])
    for _ in range(8):
        dec.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        dec.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out = {"section": name, "batch": batch, "ms_per_step": ms, "tok_s": batch * 1e3 / ms, "persistent": dec.persistent,
           "status": dec.status(), "launches": dec.launches_per_step()}
    tr = dec.read_trace() if os.environ.get('ONEBIT_PERSIST_TRACE') == '1' else None
    if tr is not None:
        L = dec.L
        for who in (0, 1):
            t = tr[who].astype(np.int64)
            lay = t[1:1 + L]
            seg = np.diff(lay[:, :6], axis=1)  # qkv, attention, o+resid, gate/up, down+resid
            sub = {"A_poll": lay[:, 6] - lay[:, 0], "A_wwait": lay[:, 20] - lay[:, 6], "A_mma": lay[:, 7] - lay[:, 20], "A_epi": lay[:, 1] - lay[:, 7],
                   "B_poll": lay[:, 17] - lay[:, 1], "B_rope": lay[:, 18] - lay[:, 17], "B_softmax": lay[:, 19] - lay[:, 18],
                   "B_out": lay[:, 2] - lay[:, 19],
                   "C_poll": lay[:, 8] - lay[:, 2], "C_wwait": lay[:, 21] - lay[:, 8], "C_mma": lay[:, 9] - lay[:, 21], "C_stats": lay[:, 10] - lay[:, 9],
                   "C_pub": lay[:, 3] - lay[:, 10],
                   "D1_poll": lay[:, 11] - lay[:, 3], "D1_wwait": lay[:, 22] - lay[:, 11], "D1_mma": lay[:, 12] - lay[:, 22], "D1_stats": lay[:, 13] - lay[:, 12],
                   "D1_pub": lay[:, 4] - lay[:, 13],
                   "D2_poll": lay[:, 14] - lay[:, 4], "D2_wwait": lay[:, 23] - lay[:, 14], "D2_mma": lay[:, 15] - lay[:, 23], "D2_stats": lay[:, 16] - lay[:, 15],
                   "D2_pub": lay[:, 5] - lay[:, 16]}
            o = {"per_layer_mean": [round(float(x) / 1e3, 2) for x in seg.mean(0)],
                 "layers_total": float(lay[-1, 5] - lay[0, 0]) / 1e3,
                 "sub_us_mean": {k: round(float(v[1:].mean()) / 1e3, 2) for k, v in sub.items()}}
            if who == 0:
                o.update(embed=float(lay[0, 0] - t[0, 0]) / 1e3, lm_head=float(t[1 + L, 1] - t[1 + L, 0]) / 1e3,
                         kernel_total=float(t[1 + L, 1] - t[0, 0]) / 1e3)
            ll = L // 2  # barrier-level trace of a middle layer: (source line, us since layer start, delta)
            n = int((t[1 + ll, 96:160] > 0).sum())
            t0 = int(t[1 + ll, 0])
            seq, prev = [], t0
            for k in range(n):
                tk = int(t[1 + ll, 32 + k])
                seq.append([int(t[1 + ll, 96 + k]), round((tk - t0) / 1e3, 2), round((tk - prev) / 1e3, 2)])
                prev = tk
            o["barriers_mid_layer"] = seq
            o["imma_C_warps"] = [[round((int(t[1 + ll, 160 + w]) - t0) / 1e3, 2), round((int(t[1 + ll, 176 + w]) - t0) / 1e3, 2)] for w in range(16)]
            o["imma_C_inner"] = [[round((int(t[1 + ll, 24 + 2 * w]) - t0) / 1e3, 2), round((int(t[1 + ll, 25 + 2 * w]) - t0) / 1e3, 2)] for w in range(4)]
            o["stage_ends_mid_layer"] = [round((int(t[1 + ll, k]) - t0) / 1e3, 2) for k in range(1, 6)]
            out[f"trace_us_cta{'0' if who == 0 else 'last'}"] = o
    dec.close()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    for arg in sys.argv[1:]:
        if arg == "tiny":
            sec_tiny()
        elif arg == "wide7b":
            wide(LLAMA_7B, "wide7b", 2, 6)
        elif arg == "wide7b_b2":
            wide(LLAMA_7B, "wide7b_b2", 2, 4, batch=2)
        elif arg == "wide13b":
            wide(LLAMA2_13B, "wide13b", 1, 4)
        elif arg == "time7b":
            timing(LLAMA_7B, "time7b")
        elif arg == "time7b_b2":
            timing(LLAMA_7B, "time7b_b2", batch=2)
        elif arg == "time13b":
            timing(LLAMA2_13B, "time13b")
        elif arg.startswith("time7b_b"):
            timing(LLAMA_7B, arg, batch=int(arg[len("time7b_b"):]), steps=32)
        elif arg.startswith("time13b_b"):
            timing(LLAMA2_13B, arg, batch=int(arg[len("time13b_b"):]), steps=32)
        else:
            raise SystemExit(f"unknown section {arg}")
