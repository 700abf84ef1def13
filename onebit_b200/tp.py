"""Tensor-parallel sharding of a OneBit LLaMA state dict (host side, pure torch/CPU).

Megatron layout: q/k/v and gate/up are column-parallel (shard the N rows of the packed sign matrix and
of `weight_scale`), o_proj and down_proj are row-parallel (shard the K byte-columns of the packed matrix and
`input_factor`). A row-parallel shard's K is zero-padded to a multiple of 256 columns (the GEMV unit): padded sign
bytes are 0 (= +1) and the padded `input_factor` is 0, so the pad contributes exactly nothing.
The LayerNorm inside every BitLinear (bitnet.py:118) spans the FULL N, hence two kinds of collective per block:
(sum, sumsq) of the column-parallel outputs and the partial sums of the row-parallel outputs.
"""
from __future__ import annotations

from typing import Dict

import torch

UNIT = 256


def pad_to(n: int, m: int = UNIT) -> int:
    return (n + m - 1) // m * m


def shard_state_dict(config: Dict, sd: Dict[str, torch.Tensor], tp: int, rank: int) -> Dict[str, torch.Tensor]:
    H, I = int(config["hidden_size"]), int(config["intermediate_size"])
    heads = int(config["num_attention_heads"])
    if heads % tp or I % tp or (I // tp) % 8:
        raise ValueError(f"cannot shard heads={heads}, intermediate={I} over tp={tp}")
    Hl, Il = H // tp, I // tp
    Hk, Ik = pad_to(Hl), pad_to(Il)
    out = {}
    for name, t in sd.items():
        if ".self_attn." in name or ".mlp." in name:
            proj = name.split(".")[-2]
            kind = name.split(".")[-1]
            col_parallel = proj in ("q_proj", "k_proj", "v_proj", "gate_proj", "up_proj")
            nl = Hl if proj.endswith(("q_proj", "k_proj", "v_proj")) else Il
            kl, kpad = (Hl, Hk) if proj == "o_proj" else (Il, Ik)
            if col_parallel:
                if kind in ("weight", "weight_scale"):
                    t = t[rank * nl:(rank + 1) * nl]
            else:  # o_proj / down_proj: shard K
                if kind == "weight":
                    t = t[:, rank * kl // 8:(rank + 1) * kl // 8]
                    t = torch.nn.functional.pad(t, (0, (kpad - kl) // 8), value=0)
                elif kind == "input_factor":
                    t = torch.nn.functional.pad(t[rank * kl:(rank + 1) * kl], (0, kpad - kl), value=0.0)
        out[name] = t.contiguous()
    return out
