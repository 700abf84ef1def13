"""Op-for-op CPU port of the reference's decode path, used ONLY as the timed CPU baseline / `--impl reference` arm of
bench.py and as a second checker in tests (TEST INFRASTRUCTURE, NOT PRODUCT — the product never imports this).

The reference itself is Python (torch ATen ops on CPU) and cannot travel to the GPU box, so this file restates the
exact op sequence it executes, per call, so that the CPU cost is the reference's cost:

  BitLinearInf.forward, transformers/src/transformers/models/bitnet.py:112-122
      x * input_factor -> unpack the int8 matrix to a dense +-1 matrix (:98-110, int64 temporaries included)
      -> F.linear -> *= weight_scale -> LayerNorm(N, no affine)
  LlamaDecoderLayerInf / LlamaAttentionInf / LlamaMLPInf / LlamaRMSNorm, modeling_bitllama.py:67-81,223-257,431-583,856-918
      eager attention with a tuple KV cache grown by torch.cat, fp32 softmax, rotate-half RoPE.

Pinned against tests/golden/tiny_model.npz (reference logits) in tests/test_ref_port_cpu.py.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F


def unpack_dense(weight_i8: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """bitnet.py:98-110: bit i of byte j -> column 8j+i, value -2*bit + 1 (arange is int64, so the shifted tensor is)."""
    shifts = torch.arange(8, device=weight_i8.device).view(1, 1, 8)
    bits = ((weight_i8.unsqueeze(-1) >> shifts) & 1).to(dtype)
    return -2 * bits.view(weight_i8.shape[0], -1) + 1


def bitlinear_forward(x: torch.Tensor, weight_i8: torch.Tensor, g: torch.Tensor, h: torch.Tensor,
                      bias: Optional[torch.Tensor] = None, eps: float = 1e-5) -> torch.Tensor:
    """bitnet.py:112-122."""
    x = x * h.view(1, -1)
    out = F.linear(x, unpack_dense(weight_i8, g.dtype))
    out *= g.view(1, -1)
    out = F.layer_norm(out, (out.shape[-1],), None, None, eps)
    if bias is not None:
        out += bias
    return out


def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """modeling_bitllama.py:67-81."""
    dt = x.dtype
    x = x.to(torch.float32)
    x = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)
    return w * x.to(dt)


def _rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


class RefPortModel:
    """Decode-step port over a reference state dict (same keys as BitLlamaForCausalLMInf)."""

    def __init__(self, config: Dict, sd: Dict[str, torch.Tensor], dtype=torch.float32):
        self.c = config
        self.sd = {k: (v if v.dtype == torch.int8 else v.to(dtype)) for k, v in sd.items()}
        self.H = int(config["hidden_size"])
        self.nh = int(config["num_attention_heads"])
        self.hd = self.H // self.nh
        self.L = int(config["num_hidden_layers"])
        self.eps = float(config.get("rms_norm_eps", 1e-6))
        theta = float(config.get("rope_theta", 10000.0))
        self.inv_freq = 1.0 / (theta ** (torch.arange(0, self.hd, 2).float() / self.hd))
        self.dtype = dtype

    def _bl(self, name: str, x: torch.Tensor) -> torch.Tensor:
        return bitlinear_forward(x, self.sd[name + ".weight"], self.sd[name + ".weight_scale"],
                                 self.sd[name + ".input_factor"])

    def layer(self, l: int, hs: torch.Tensor, pos: int, past: Optional[Tuple[torch.Tensor, torch.Tensor]]):
        """One LlamaDecoderLayerInf.forward for q_len tokens starting at `pos` (:869-930)."""
        pre = f"model.layers.{l}."
        b, q_len, _ = hs.shape
        res = hs
        x = rms_norm(hs, self.sd[pre + "input_layernorm.weight"], self.eps)
        q = self._bl(pre + "self_attn.q_proj", x).view(b, q_len, self.nh, self.hd).transpose(1, 2)
        k = self._bl(pre + "self_attn.k_proj", x).view(b, q_len, self.nh, self.hd).transpose(1, 2)
        v = self._bl(pre + "self_attn.v_proj", x).view(b, q_len, self.nh, self.hd).transpose(1, 2)
        t = torch.arange(pos, pos + q_len, dtype=self.inv_freq.dtype)
        freqs = torch.einsum("i,j->ij", t, self.inv_freq)
        emb = torch.cat((freqs, freqs), dim=-1)
        cos, sin = emb.cos().to(q.dtype)[None, None], emb.sin().to(q.dtype)[None, None]
        q = q * cos + _rotate_half(q) * sin
        k = k * cos + _rotate_half(k) * sin
        if past is not None:
            k = torch.cat([past[0], k], dim=2)
            v = torch.cat([past[1], v], dim=2)
        att = torch.matmul(q, k.transpose(2, 3)) / math.sqrt(self.hd)
        kv_len = k.shape[2]
        if q_len > 1:
            mask = torch.full((q_len, kv_len), torch.finfo(att.dtype).min)
            mask = torch.triu(mask, diagonal=kv_len - q_len + 1)
            att = att + mask
        att = F.softmax(att, dim=-1, dtype=torch.float32).to(q.dtype)
        o = torch.matmul(att, v).transpose(1, 2).reshape(b, q_len, self.H)
        hs = res + self._bl(pre + "self_attn.o_proj", o)
        res = hs
        x = rms_norm(hs, self.sd[pre + "post_attention_layernorm.weight"], self.eps)
        gate = self._bl(pre + "mlp.gate_proj", x)
        up = self._bl(pre + "mlp.up_proj", x)
        hs = res + self._bl(pre + "mlp.down_proj", F.silu(gate) * up)
        return hs, (k, v)

    def forward(self, ids: torch.Tensor, past: Optional[List] = None, pos: int = 0):
        hs = F.embedding(ids, self.sd["model.embed_tokens.weight"])
        new_past = []
        for l in range(self.L):
            hs, kv = self.layer(l, hs, pos, past[l] if past is not None else None)
            new_past.append(kv)
        hs = rms_norm(hs, self.sd["model.norm.weight"], self.eps)
        return F.linear(hs, self.sd["lm_head.weight"]).float(), new_past
