"""`BitLinearB200` — the nn.Module face of the CUDA path, interface-identical to the reference's
`BitLinearInf` (transformers/src/transformers/models/bitnet.py:71-122):

    BitLinearInf(in_features, out_features, groups=1, bias=False, device=None, dtype=None)
    parameters: weight int8 [N, K/8], weight_scale [N], input_factor [K], bias [N] | None   (all frozen)
    forward(input[..., K]) -> [..., N] = LayerNorm_N(weight_scale * (sign(W) @ (input_factor * input))) (+ bias)

State-dict keys, dtypes and shapes are the reference's, so `from_pretrained` / `save_pretrained` checkpoints
round-trip unchanged. The arithmetic runs in libonebit_b200.so (hand-written sm_100a kernels); there is no
PyTorch or CPU fallback — a CPU tensor, a missing library or a non-B200 device raises RuntimeError.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch import nn

from . import _lib

_DTYPE_CODE = {torch.float16: _lib.F16, torch.bfloat16: _lib.BF16, torch.float32: _lib.F32}


def _code(dt: torch.dtype, what: str) -> int:
    try:
        return _DTYPE_CODE[dt]
    except KeyError:
        raise RuntimeError(f"onebit_b200: unsupported {what} dtype {dt} (supported: float16, bfloat16, float32)")


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"onebit_b200: `{name}` lives on {t.device}; the 1-bit linear path is CUDA (sm_100a) only and has no "
            "CPU fallback. Move the module and its input to a B200 device.")


def _check_layer_args(x, weight, weight_scale, input_factor, bias):
    _require_cuda(x, "input")
    for name, p in (("weight", weight), ("weight_scale", weight_scale), ("input_factor", input_factor)):
        if p.device != x.device:
            raise RuntimeError(f"onebit_b200: `{name}` is on {p.device} but the input is on {x.device}")
    if weight.dtype != torch.int8:
        raise RuntimeError(f"onebit_b200: `weight` must be int8 bit-packed signs, got {weight.dtype}")
    if weight.dim() != 2 or not weight.is_contiguous():
        raise RuntimeError("onebit_b200: `weight` must be a contiguous [out_features, in_features // 8] tensor")
    n, kb = weight.shape
    k = kb * 8
    if x.shape[-1] != k:
        raise RuntimeError(f"onebit_b200: input has {x.shape[-1]} features, the packed weight expects {k}")
    if input_factor.numel() != k or weight_scale.numel() != n:
        raise RuntimeError("onebit_b200: weight_scale / input_factor sizes do not match the packed weight")
    if weight_scale.dtype != input_factor.dtype or (bias is not None and bias.dtype != weight_scale.dtype):
        raise RuntimeError("onebit_b200: weight_scale, input_factor and bias must share one dtype")
    return n, k


def bitlinear_matvec(x, weight, weight_scale, input_factor, scale_by_g: bool = True, variant: str = "auto"):
    """t = sign(W) @ (input_factor * x) (optionally * weight_scale), fp32 [..., N]  (bitnet.py:113-116)."""
    n, k = _check_layer_args(x, weight, weight_scale, input_factor, None)
    lib = _lib.load()
    x2 = x.reshape(-1, k).contiguous()
    m = x2.shape[0]
    t = torch.empty((m, n), dtype=torch.float32, device=x.device)
    g = weight_scale.contiguous()
    h = input_factor.contiguous()
    ws_bytes = lib.onebit_matvec_workspace_bytes(m, k)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.onebit_bitlinear_matvec(x2.data_ptr(), weight.data_ptr(), g.data_ptr(), h.data_ptr(), t.data_ptr(),
                                         m, k, n, _code(x.dtype, "activation"), _code(g.dtype, "parameter"),
                                         int(scale_by_g), ws.data_ptr(), ws_bytes, _lib.VARIANTS[variant], _stream(x))
    _lib.check(rc, "onebit_bitlinear_matvec")
    return t.reshape(*x.shape[:-1], n)


def scale_layernorm(t, weight_scale, bias, out_dtype: torch.dtype, eps: float = 1e-5):
    """y = LayerNorm_N(weight_scale * t) (+ bias)  (bitnet.py:116-120); weight_scale=None -> t already scaled."""
    _require_cuda(t, "t")
    if t.dtype != torch.float32:
        raise RuntimeError("onebit_b200: scale_layernorm expects fp32 `t`")
    n = t.shape[-1]
    t2 = t.reshape(-1, n).contiguous()
    m = t2.shape[0]
    y = torch.empty((m, n), dtype=out_dtype, device=t.device)
    pd = weight_scale.dtype if weight_scale is not None else (bias.dtype if bias is not None else torch.float32)
    with torch.cuda.device(t.device):
        rc = _lib.load().onebit_scale_layernorm(
            t2.data_ptr(), weight_scale.data_ptr() if weight_scale is not None else None,
            bias.data_ptr() if bias is not None else None, y.data_ptr(), m, n, _code(out_dtype, "activation"),
            _code(pd, "parameter"), float(eps), _stream(t))
    _lib.check(rc, "onebit_scale_layernorm")
    return y.reshape(*t.shape[:-1], n)


def bitlinear_forward(x: torch.Tensor, weight: torch.Tensor, weight_scale: torch.Tensor, input_factor: torch.Tensor,
                      bias: Optional[torch.Tensor] = None, eps: float = 1e-5, variant: str = "auto") -> torch.Tensor:
    """Functional form of BitLinearInf.forward (bitnet.py:112-122) on the CUDA path. Returns a fresh tensor with
    the dtype/device of `x`."""
    if variant == "auto":
        # the torch dispatcher op over the same C ABI (csrc_torch/torch_ops.cpp): argument checks, output and scratch
        # allocation and the launch happen in C++ — no ctypes marshalling on the module's forward path
        return _lib.torch_ops().bitlinear(x, weight, weight_scale, input_factor, bias, float(eps))
    n, k = _check_layer_args(x, weight, weight_scale, input_factor, bias)
    lib = _lib.load()
    x2 = x.reshape(-1, k)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    m = x2.shape[0]
    y = torch.empty((m, n), dtype=x.dtype, device=x.device)
    if m == 0:
        return y.reshape(*x.shape[:-1], n)
    g = weight_scale.contiguous()
    h = input_factor.contiguous()
    b = bias.contiguous() if bias is not None else None
    ws_bytes = lib.onebit_bitlinear_workspace_bytes(m, k, n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)  # caching allocator: graph-capture safe
    with torch.cuda.device(x.device):
        rc = lib.onebit_bitlinear_forward(x2.data_ptr(), weight.data_ptr(), g.data_ptr(), h.data_ptr(),
                                          b.data_ptr() if b is not None else None, y.data_ptr(), m, k, n,
                                          _code(x.dtype, "activation"), _code(g.dtype, "parameter"), float(eps),
                                          ws.data_ptr(), ws_bytes, _lib.VARIANTS[variant], _stream(x))
    _lib.check(rc, "onebit_bitlinear_forward")
    return y.reshape(*x.shape[:-1], n)


def pack_signs(w: torch.Tensor) -> torch.Tensor:
    """GPU version of fp16_to_int8 (scripts/convert_llama_to_infer_ckpt.py:7-15): [N, K] of +-1 (0 counts as
    +1) -> int8 [N, K/8], column 8j+i in bit i of byte j, bit 1 <=> -1."""
    _require_cuda(w, "w")
    if w.dim() != 2 or w.shape[1] % 8 != 0:
        raise RuntimeError("onebit_b200: pack_signs expects [N, K] with K % 8 == 0")
    w = w.contiguous()
    out = torch.empty((w.shape[0], w.shape[1] // 8), dtype=torch.int8, device=w.device)
    with torch.cuda.device(w.device):
        rc = _lib.load().onebit_pack_signs(w.data_ptr(), out.data_ptr(), w.shape[0], w.shape[1],
                                           _code(w.dtype, "weight"), _stream(w))
    _lib.check(rc, "onebit_pack_signs")
    return out


def unpack_signs(packed: torch.Tensor, dtype: torch.dtype = torch.float16) -> torch.Tensor:
    """GPU version of BitLinearInf.int8_to_fp16 (bitnet.py:98-110): int8 [N, K/8] -> +-1 [N, K] in `dtype`."""
    _require_cuda(packed, "packed")
    if packed.dtype != torch.int8 or packed.dim() != 2:
        raise RuntimeError("onebit_b200: unpack_signs expects an int8 [N, K/8] tensor")
    packed = packed.contiguous()
    out = torch.empty((packed.shape[0], packed.shape[1] * 8), dtype=dtype, device=packed.device)
    with torch.cuda.device(packed.device):
        rc = _lib.load().onebit_unpack_signs(packed.data_ptr(), out.data_ptr(), packed.shape[0], packed.shape[1] * 8,
                                             _code(dtype, "output"), _stream(packed))
    _lib.check(rc, "onebit_unpack_signs")
    return out


class BitLinearB200(nn.Module):
    """Drop-in for the reference's `BitLinearInf` (bitnet.py:71-122). Same constructor, same parameter names,
    dtypes and shapes, same forward contract."""

    def __init__(self, in_features, out_features, groups=1, bias=False, device=None, dtype=None):
        super().__init__()
        if in_features % 8 != 0:
            raise ValueError(f"in_features must be a multiple of 8 (8 sign bits per int8 byte), got {in_features}")
        factory_kwargs = {"device": device, "dtype": dtype}
        self.in_features = in_features
        self.out_features = out_features
        self.groups = groups  # unused, as in the reference (bitnet.py:77)
        self.variant = "auto"
        self.weight = nn.Parameter(torch.empty((out_features, in_features // 8), device=device, dtype=torch.int8),
                                   requires_grad=False)
        self.weight_scale = nn.Parameter(torch.empty(out_features, **factory_kwargs), requires_grad=False)
        self.input_factor = nn.Parameter(torch.empty(in_features, **factory_kwargs), requires_grad=False)
        if bias:
            self.bias = nn.Parameter(torch.empty(out_features, **factory_kwargs), requires_grad=False)
        else:
            self.register_parameter("bias", None)
        # bitnet.py:86 — kept as an attribute like the reference's (no parameters, no state-dict keys); the fused kernels
        # read its eps, they never call it
        self.layernorm = nn.LayerNorm(out_features, elementwise_affine=False)
        self.reset_parameters()

    def reset_parameters(self):
        # bitnet.py:89-96: g = h = 1, packed weight = 0 (all signs +1), bias ~ U(+-1/sqrt(fan_in)), fan_in = K/8
        with torch.no_grad():
            self.weight_scale.fill_(1.0)
            self.input_factor.fill_(1.0)
            self.weight.zero_()
            if self.bias is not None:
                fan_in = self.weight.shape[1]
                bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
                self.bias.uniform_(-bound, bound)

    @classmethod
    def from_reference(cls, ref: nn.Module) -> "BitLinearB200":
        """Adopt the parameters of a reference `BitLinearInf` (same Parameter objects: no copy, state_dict and
        `.to()` keep working on the parent model)."""
        new = cls.__new__(cls)
        nn.Module.__init__(new)
        new.in_features = ref.in_features
        new.out_features = ref.out_features
        new.groups = getattr(ref, "groups", 1)
        ln = getattr(ref, "layernorm", None)
        new.layernorm = ln if isinstance(ln, nn.LayerNorm) else nn.LayerNorm(ref.out_features, elementwise_affine=False)
        new.variant = "auto"
        new.weight = ref.weight
        new.weight_scale = ref.weight_scale
        new.input_factor = ref.input_factor
        if getattr(ref, "bias", None) is not None:
            new.bias = ref.bias
        else:
            new.register_parameter("bias", None)
        new.train(ref.training)
        return new

    @property
    def eps(self) -> float:
        return float(self.layernorm.eps)

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        return bitlinear_forward(input, self.weight, self.weight_scale, self.input_factor, self.bias, self.eps,
                                 self.variant)

    def extra_repr(self) -> str:
        return (f"in_features={self.in_features}, out_features={self.out_features}, "
                f"bias={self.bias is not None}, packed=1bit")


def _is_reference_bitlinear_inf(mod: nn.Module) -> bool:
    if isinstance(mod, BitLinearB200):
        return False
    w = getattr(mod, "weight", None)
    return (type(mod).__name__ == "BitLinearInf" and isinstance(w, torch.Tensor) and w.dtype == torch.int8
            and hasattr(mod, "weight_scale") and hasattr(mod, "input_factor"))


def replace_bitlinear(model: nn.Module) -> int:
    """Swap every reference `BitLinearInf` inside `model` (e.g. a `BitLlamaForCausalLMInf`, whose q/k/v/o and
    gate/up/down projections are created at modeling_bitllama.py:229-231,451-454) for a `BitLinearB200` that
    shares its parameters. Returns the number of modules replaced."""
    count = 0
    for parent in list(model.modules()):
        for name, child in list(parent.named_children()):
            if _is_reference_bitlinear_inf(child):
                setattr(parent, name, BitLinearB200.from_reference(child))
                count += 1
    return count
