"""Host-side mirror of the reference's `BitLlamaForCausalLMInf` decode path over the fused CUDA decode step
(libonebit_b200.so: onebit_decoder_*).

Reference being mirrored (xuyuzhuang11/OneBit, transformers/src/transformers/models/bitllama/modeling_bitllama.py):
`BitLlamaForCausalLMInf` :1512 -> `LlamaModelInf` :1189 -> `LlamaDecoderLayerInf` :856 -> `LlamaAttentionInf` :431 /
`LlamaMLPInf` :223, and the greedy loop of generation/utils.py:2491-2571. The state dict uses the reference's keys
unchanged (`model.layers.N.self_attn.q_proj.weight|weight_scale|input_factor`, ...), so a checkpoint produced by
scripts/convert_llama_to_infer_ckpt.py loads as is.

There is no PyTorch fallback: every arithmetic step runs in hand-written sm_100a kernels.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib

_PROJ = (("q", "self_attn.q_proj"), ("k", "self_attn.k_proj"), ("v", "self_attn.v_proj"), ("o", "self_attn.o_proj"),
         ("gate", "mlp.gate_proj"), ("up", "mlp.up_proj"), ("down", "mlp.down_proj"))


def rope_tables(head_dim: int, max_seq_len: int, theta: float = 10000.0, device="cpu"):
    """cos/sin exactly as LlamaRotaryEmbedding builds them (modeling_bitllama.py:94-111), fp32, first half only
    (the reference concatenates two identical halves)."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2).float() / head_dim))
    t = torch.arange(max_seq_len, dtype=inv_freq.dtype)
    freqs = torch.einsum("i,j->ij", t, inv_freq)
    return freqs.cos().contiguous().to(device), freqs.sin().contiguous().to(device)


class BitLlamaDecoderB200:
    """Greedy decoder for a OneBit LLaMA (`BitLlamaForCausalLMInf`) on one B200.

    config keys used: hidden_size, intermediate_size, num_hidden_layers, num_attention_heads, vocab_size,
    rms_norm_eps, rope_theta (names of BitLlamaConfig, configuration_bitllama.py:115-136).
    """

    def __init__(self, config: Dict, state_dict: Dict[str, torch.Tensor], device="cuda:0", max_seq_len: int = 2048,
                 max_batch: int = 1, param_dtype: torch.dtype = torch.float16, use_graph: bool = True,
                 tp_group=None):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("onebit_b200: the decoder is CUDA (sm_100a) only; there is no CPU fallback")
        self.config = dict(config)
        self.H = int(config["hidden_size"])
        self.I = int(config["intermediate_size"])
        self.L = int(config["num_hidden_layers"])
        self.n_heads = int(config["num_attention_heads"])
        self.V = int(config["vocab_size"])
        if int(config.get("num_key_value_heads", self.n_heads)) != self.n_heads:
            raise RuntimeError("onebit_b200: grouped-query attention is not built (LLaMA-7B/13B use MHA)")
        self.max_seq_len = int(max_seq_len)
        self.max_batch = int(max_batch)
        self.param_dtype = param_dtype
        self.use_graph = use_graph
        pcode = {torch.float16: _lib.F16, torch.bfloat16: _lib.BF16, torch.float32: _lib.F32}[param_dtype]

        dev = self.device
        self._keep = []  # device tensors the C side points into

        def put(t: torch.Tensor, dtype=None) -> torch.Tensor:
            t = t.detach().to(dev, dtype if dtype is not None else t.dtype).contiguous()
            self._keep.append(t)
            return t

        # tensor parallelism: shard the state dict for this rank (o/down K shards zero-padded), all-reduce through
        # torch.distributed on the current stream (NCCL; capturable in the step's CUDA graph)
        self.tp_size, self.tp_rank, self._ar_cb = 1, 0, None
        if tp_group is not None:
            import torch.distributed as dist
            from .tp import shard_state_dict
            self.tp_size, self.tp_rank = dist.get_world_size(tp_group), dist.get_rank(tp_group)
            if self.tp_size > 1:
                state_dict = shard_state_dict(config, state_dict, self.tp_size, self.tp_rank)

                def _allreduce(user, ptr, count, stream):
                    try:
                        t = _tensor_from_ptr(ptr, (int(count),), torch.float32, dev)
                        dist.all_reduce(t, group=tp_group)
                        return 0
                    except Exception as exc:  # surfaced as an error code by the C side
                        print(f"onebit_b200: all-reduce failed: {exc}", flush=True)
                        return 1

                self._ar_cb = _lib.ALLREDUCE_FN(_allreduce)
        sd = state_dict
        layers = (_lib.LayerParams * self.L)()
        self.weight_bytes = 0
        for l in range(self.L):
            pre = f"model.layers.{l}."
            for field, name in _PROJ:
                w = sd[pre + name + ".weight"]
                if w.dtype != torch.int8:
                    raise RuntimeError(f"{pre + name}.weight must be int8 bit-packed signs (BitLinearInf), got {w.dtype}")
                w = put(w)
                g = put(sd[pre + name + ".weight_scale"], param_dtype)
                h = put(sd[pre + name + ".input_factor"], param_dtype)
                bp = getattr(layers[l], field)
                bp.weight, bp.weight_scale, bp.input_factor = w.data_ptr(), g.data_ptr(), h.data_ptr()
                self.weight_bytes += w.numel()
            layers[l].input_layernorm = put(sd[pre + "input_layernorm.weight"], param_dtype).data_ptr()
            layers[l].post_attention_layernorm = put(sd[pre + "post_attention_layernorm.weight"], param_dtype).data_ptr()
        embed = put(sd["model.embed_tokens.weight"], torch.float16)
        final_norm = put(sd["model.norm.weight"], param_dtype)
        lm_head = put(sd["lm_head.weight"], torch.float16)
        cos, sin = rope_tables(self.H // self.n_heads, self.max_seq_len, float(config.get("rope_theta", 10000.0)))
        cos, sin = put(cos), put(sin)
        cfg = _lib.DecoderConfig(self.H, self.I, self.L, self.n_heads, self.V, self.max_seq_len, self.max_batch, pcode,
                                 float(config.get("rms_norm_eps", 1e-6)), 1e-5, self.tp_size, self.tp_rank)
        self._layers = layers
        self._handle = ctypes.c_void_p()
        with torch.cuda.device(dev):
            rc = self.lib.onebit_decoder_create(ctypes.byref(self._handle), ctypes.byref(cfg), layers, embed.data_ptr(),
                                                final_norm.data_ptr(), lm_head.data_ptr(), cos.data_ptr(), sin.data_ptr(),
                                                self._ar_cb if self._ar_cb is not None else _lib.ALLREDUCE_FN(0), None)
        _lib.check(rc, "onebit_decoder_create")
        self.tp_allreduce = "none" if self.tp_size == 1 else "nccl"
        if self.tp_size > 1 and os.environ.get("ONEBIT_TP_ALLREDUCE", "p2p") != "nccl":
            self._enable_p2p_allreduce(tp_group)
        self.logits = torch.zeros((self.max_batch, self.V), dtype=torch.float32, device=dev)
        self.forced = torch.zeros((self.max_batch,), dtype=torch.int64, device=dev)
        self._graphs = {}
        self._warmed = set()
        self.batch = 0
        self._pos_hi = 0  # upper bound of every sequence's position (the C side keeps the same count for eager calls)
        self.persistent = bool(self.lib.onebit_decoder_is_persistent(self._handle))

    def _enable_p2p_allreduce(self, tp_group):
        """One-shot all-reduce over NVLink peer memory (csrc/p2p_allreduce.cu) instead of NCCL: a symmetric buffer per rank
        (torch symmetric memory maps every peer's buffer into this process), filled with the "not written" pattern -0.0f.
        Falls back to the NCCL callback, loudly, if symmetric memory cannot be set up on this system."""
        import torch.distributed as dist
        try:
            import torch.distributed._symmetric_memory as symm
            nfloats = 3 * self.tp_size * max(self.max_batch * self.H, 64)
            buf = symm.empty(nfloats, dtype=torch.float32, device=self.device)
            buf.fill_(-0.0)
            hdl = symm.rendezvous(buf, tp_group)
            torch.cuda.synchronize(self.device)
            dist.barrier(group=tp_group)
            ptrs = (ctypes.c_void_p * self.tp_size)(*[int(p) for p in hdl.buffer_ptrs])
            rc = self.lib.onebit_decoder_enable_p2p_allreduce(self._handle, self.tp_rank, self.tp_size, ptrs, nfloats * 4)
            _lib.check(rc, "onebit_decoder_enable_p2p_allreduce")
            self._symm = (buf, hdl)
            self.tp_allreduce = "p2p"
        except Exception as exc:
            print(f"onebit_b200: one-shot NVLink all-reduce unavailable ({type(exc).__name__}: {exc}); using NCCL", flush=True)

    # ------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle:
            self.lib.onebit_decoder_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _set_state(self, ids: torch.Tensor, pos: torch.Tensor):
        with torch.cuda.device(self.device):
            rc = self.lib.onebit_decoder_reset(self._handle, ids.data_ptr(), pos.data_ptr(), ids.numel(), self._stream())
        _lib.check(rc, "onebit_decoder_reset")

    def reset(self, first_ids, positions=None):
        """Start `batch` sequences: ids to feed at the next step and their positions (default 0)."""
        ids = torch.as_tensor(first_ids, dtype=torch.int64).reshape(-1).cpu().contiguous()
        b = ids.numel()
        pos = torch.zeros(b, dtype=torch.int32) if positions is None else \
            torch.as_tensor(positions, dtype=torch.int32).reshape(-1).cpu().contiguous()
        if int(pos.max()) >= self.max_seq_len or int(pos.min()) < 0:
            raise RuntimeError(f"onebit_b200: positions must lie in [0, max_seq_len={self.max_seq_len})")
        if int(ids.min()) < 0 or int(ids.max()) >= self.V:
            raise RuntimeError(f"onebit_b200: token ids must lie in [0, vocab_size={self.V})")
        self.batch = b
        if b not in self._warmed:
            # one eager step per batch size before any graph capture: kernel attributes / lazy module loading
            # must not happen inside a capture. The state it scribbles is overwritten right below.
            self._set_state(torch.zeros(b, dtype=torch.int64), torch.zeros(b, dtype=torch.int32))
            with torch.cuda.device(self.device):
                self._enqueue(True)
                self._enqueue(False)
            torch.cuda.synchronize(self.device)
            self._warmed.add(b)
        self._set_state(ids, pos)
        self._pos_hi = int(pos.max())

    def _enqueue(self, forced: bool):
        rc = self.lib.onebit_decoder_step(self._handle, self.batch, self.forced.data_ptr() if forced else None,
                                          self.logits.data_ptr(), self._stream())
        _lib.check(rc, "onebit_decoder_step")

    def step(self, forced_ids: Optional[torch.Tensor] = None):
        """One decode step for all sequences. Feeds `forced_ids` if given, else the previous step's argmax.
        Leaves logits in `self.logits[:batch]` and the next ids on the device (`next_ids()`)."""
        forced = forced_ids is not None
        if self._pos_hi >= self.max_seq_len:
            raise RuntimeError(f"onebit_b200: a sequence has reached max_seq_len={self.max_seq_len}; the static KV cache is "
                               "full (create the decoder with a larger max_seq_len)")
        self._pos_hi += 1
        if forced:
            self.forced[: self.batch].copy_(forced_ids.reshape(-1), non_blocking=True)  # H2D if the ids are on the host
        with torch.cuda.device(self.device):
            if not self.use_graph:
                self._enqueue(forced)
                return
            key = (self.batch, forced)
            g = self._graphs.get(key)
            if g is None:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):  # capture only records; state is untouched until replay
                    self._enqueue(forced)
                self._graphs[key] = g
            g.replay()

    def next_ids(self) -> torch.Tensor:
        ptr = self.lib.onebit_decoder_next_ids(self._handle)
        return _tensor_from_ptr(ptr, (self.max_batch,), torch.int64, self.device)[: self.batch]

    def positions(self) -> torch.Tensor:
        ptr = self.lib.onebit_decoder_positions(self._handle)
        return _tensor_from_ptr(ptr, (self.max_batch,), torch.int32, self.device)[: self.batch]

    def launches_per_step(self) -> int:
        return int(self.lib.onebit_decoder_kernel_launches_per_step(self._handle))

    def status(self) -> int:
        """0 = fine; 1 = an in-kernel exchange of the persistent step timed out; 2 = decode past max_seq_len refused."""
        code = ctypes.c_int(0)
        _lib.check(self.lib.onebit_decoder_status(self._handle, ctypes.byref(code)), "onebit_decoder_status")
        return int(code.value)

    def read_trace(self) -> Optional[np.ndarray]:
        """Stage time stamps (ns) of the last persistent step: array [2 tracer CTAs, L + 2, 192] (see onebit_b200.h),
        or None."""
        if not self.persistent:
            return None
        n = 2 * 192 * (self.L + 2)
        buf = np.zeros(n, dtype=np.uint64)
        got = self.lib.onebit_decoder_read_trace(self._handle, buf.ctypes.data, n)
        if got <= 0:
            return None
        return buf.reshape(2, self.L + 2, 192)

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def prefill(self, input_ids: torch.Tensor, all_logits: bool = False, pos0: int = 0) -> torch.Tensor:
        """Prompt pass over [B, T] token ids in ONE go (the reference's q_len > 1 forward, modeling_bitllama.py:1217-1315):
        every BitLinear is a tcgen05 GEMM over B*T tokens, attention is causal flash attention, K/V go to the static
        cache. Returns the last-token logits [B, V] (or all logits [B, T, V] with `all_logits`); afterwards `step()` /
        `next_ids()` continue the sequences (next ids = greedy continuation of each prompt)."""
        if self.tp_size > 1:
            raise RuntimeError("onebit_b200: prefill() is single-GPU; tensor-parallel decoders feed the prompt with step()")
        input_ids = torch.as_tensor(input_ids, dtype=torch.int64)
        b, t = input_ids.shape
        if b > self.max_batch or pos0 + t > self.max_seq_len:
            raise RuntimeError(f"onebit_b200: prompt [{b}, {t}] at position {pos0} does not fit max_batch={self.max_batch}, "
                               f"max_seq_len={self.max_seq_len}")
        if int(input_ids.min()) < 0 or int(input_ids.max()) >= self.V:
            raise RuntimeError(f"onebit_b200: token ids must lie in [0, vocab_size={self.V})")
        if b not in self._warmed:
            self.reset(input_ids[:, 0])  # eager warm-up steps of the decode kernels (they must not load inside a graph capture)
        ids_dev = input_ids.to(self.device).contiguous()
        last = torch.empty((b, self.V), dtype=torch.float32, device=self.device)
        full = torch.empty((b * t, self.V), dtype=torch.float32, device=self.device) if all_logits else None
        with torch.cuda.device(self.device):
            rc = self.lib.onebit_decoder_prefill(self._handle, b, t, int(pos0), ids_dev.data_ptr(), last.data_ptr(),
                                                 full.data_ptr() if full is not None else None, self._stream())
        _lib.check(rc, "onebit_decoder_prefill")
        self.batch = b
        self._pos_hi = pos0 + t
        self.logits[:b].copy_(last)
        return full.view(b, t, self.V) if all_logits else last

    @torch.no_grad()
    def forward_tokens(self, input_ids: torch.Tensor) -> torch.Tensor:
        """Teacher-forced pass over [B, T] token ids through the decode path; returns logits [B, T, V] (fp32) —
        the quantity BitLlamaForCausalLMInf.forward returns (:1546-1611), computed one position at a time."""
        input_ids = torch.as_tensor(input_ids, dtype=torch.int64)
        b, t = input_ids.shape
        if t > self.max_seq_len:
            raise RuntimeError(f"onebit_b200: {t} tokens do not fit max_seq_len={self.max_seq_len}")
        self.reset(input_ids[:, 0])
        out = torch.empty((b, t, self.V), dtype=torch.float32, device=self.device)
        ids_dev = input_ids.to(self.device)
        for i in range(t):
            self.step(ids_dev[:, i])
            out[:, i].copy_(self.logits[:b])
        return out

    @torch.no_grad()
    def generate(self, prompt_ids: torch.Tensor, max_new_tokens: int, use_prefill: bool = False) -> torch.Tensor:
        """Greedy decoding (generation/utils.py:2491-2571 with do_sample=False, no EOS stop): returns
        [B, T0 + max_new_tokens] like `model.generate`."""
        prompt_ids = torch.as_tensor(prompt_ids, dtype=torch.int64)
        b, t0 = prompt_ids.shape
        if t0 + max_new_tokens - 1 > self.max_seq_len:
            raise RuntimeError(f"onebit_b200: prompt {t0} + {max_new_tokens} new tokens do not fit max_seq_len={self.max_seq_len}")
        ids_dev = prompt_ids.to(self.device)
        if use_prefill:  # the whole prompt in one pass (tcgen05 GEMMs + causal flash attention), then decode steps
            self.prefill(prompt_ids)
        else:
            self.reset(prompt_ids[:, 0])
            for i in range(t0):
                self.step(ids_dev[:, i])
        new = [self.next_ids().clone()]
        for _ in range(max_new_tokens - 1):
            self.step()
            new.append(self.next_ids().clone())
        return torch.cat([ids_dev, torch.stack(new, dim=1)], dim=1)

    def perplexity(self, input_ids: torch.Tensor, use_prefill: bool = False) -> float:
        """evaluation/lm_eval.py:99-124: per window CE(mean over shifted tokens) * seqlen, exp(sum / (n * seqlen))."""
        logits = self.prefill(input_ids, all_logits=True) if use_prefill else self.forward_tokens(input_ids)
        b, t, _ = logits.shape
        ids = torch.as_tensor(input_ids).to(self.device)
        nll = 0.0
        for i in range(b):
            loss = torch.nn.functional.cross_entropy(logits[i, :-1].double(), ids[i, 1:])
            nll += float(loss) * t
        return float(np.exp(nll / (b * t)))


def _tensor_from_ptr(ptr: int, shape, dtype, device) -> torch.Tensor:
    """Wrap decoder-owned device memory as a torch tensor (no copy) through the CUDA array interface."""
    n = int(np.prod(shape))
    itemsize = torch.empty((), dtype=dtype).element_size()
    typestr = {torch.int64: "<i8", torch.int32: "<i4", torch.float32: "<f4"}[dtype]

    class _Holder:
        __cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3,
                                    "strides": None}
    del n, itemsize
    return torch.as_tensor(_Holder(), device=device)


def synthetic_state_dict(config: Dict, seed: int = 0, param_dtype=torch.float16) -> Dict[str, torch.Tensor]:
    """Random-init OneBit LLaMA weights with the reference's state-dict keys (SURVEY.md §8d recipe): random packed
    bytes, g ~ U(0.5, 1.5), h ~ U(-1.5, 1.5), norm weights ~ U(0.5, 1.5), embeddings / lm_head ~ N(0, 0.02)."""
    gen = torch.Generator().manual_seed(seed)
    H, I, L, V = (int(config[k]) for k in ("hidden_size", "intermediate_size", "num_hidden_layers", "vocab_size"))
    sd = {}
    shapes = {"self_attn.q_proj": (H, H), "self_attn.k_proj": (H, H), "self_attn.v_proj": (H, H),
              "self_attn.o_proj": (H, H), "mlp.gate_proj": (I, H), "mlp.up_proj": (I, H), "mlp.down_proj": (H, I)}
    for l in range(L):
        pre = f"model.layers.{l}."
        for name, (n, k) in shapes.items():
            sd[pre + name + ".weight"] = torch.randint(-128, 128, (n, k // 8), dtype=torch.int8, generator=gen)
            sd[pre + name + ".weight_scale"] = (torch.rand(n, generator=gen) + 0.5).to(param_dtype)
            sd[pre + name + ".input_factor"] = (torch.rand(k, generator=gen) * 3 - 1.5).to(param_dtype)
        sd[pre + "input_layernorm.weight"] = (torch.rand(H, generator=gen) + 0.5).to(param_dtype)
        sd[pre + "post_attention_layernorm.weight"] = (torch.rand(H, generator=gen) + 0.5).to(param_dtype)
    sd["model.embed_tokens.weight"] = (torch.randn(V, H, generator=gen) * 0.02).half()
    sd["model.norm.weight"] = (torch.rand(H, generator=gen) + 0.5).to(param_dtype)
    sd["lm_head.weight"] = (torch.randn(V, H, generator=gen) * 0.02).half()
    return sd


LLAMA_7B = dict(hidden_size=4096, intermediate_size=11008, num_hidden_layers=32, num_attention_heads=32,
                vocab_size=32000, rms_norm_eps=1e-6, rope_theta=10000.0)
LLAMA2_13B = dict(hidden_size=5120, intermediate_size=13824, num_hidden_layers=40, num_attention_heads=40,
                  vocab_size=32000, rms_norm_eps=1e-5, rope_theta=10000.0)
