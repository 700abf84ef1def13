"""CPU oracle for the OneBit 1-bit linear layer — TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference`` legs may
import this module. The product (``onebit_b200``) never does; it fails loudly without its CUDA library.

Two restatements of the reference algorithm (xuyuzhuang11/OneBit @ 42d6d7b) live here:

* numpy (this file)  — readable, used for small cases and to cross-check the C one;
* C (onebit_oracle.c, loaded through ctypes) — fast enough for the full LLaMA shapes.

Reference lines restated:
  pack    scripts/convert_llama_to_infer_ckpt.py:7-15   bit = (1 - sign)/2, column 8j+i -> bit i of byte j
  unpack  transformers/src/transformers/models/bitnet.py:98-110
  forward transformers/src/transformers/models/bitnet.py:112-122  y = LN_N(g * (S @ (h*x))) (+ bias)

Pinning status: PINNED — `tests/test_oracle_golden.py` checks both restatements against
tests/golden/*.npz, which hold outputs of the reference's own code executed in the build container
(tests/golden/gen_golden.py). The reference itself ships no tests or golden vectors.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libonebit_oracle.so"
_lib = None


def build(force: bool = False) -> Path:
    """Compile onebit_oracle.c next to its source (gcc, a second or two)."""
    src = _HERE / "onebit_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-s"], check=True, env={**os.environ, "MAKEFLAGS": ""})
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(str(_LIB_PATH))
        f32p = ctypes.POINTER(ctypes.c_float)
        i8p = ctypes.POINTER(ctypes.c_int8)
        i64 = ctypes.c_int64
        lib.onebit_oracle_pack.argtypes = [f32p, i8p, i64, i64]
        lib.onebit_oracle_unpack.argtypes = [i8p, f32p, i64, i64]
        lib.onebit_oracle_forward.argtypes = [f32p, i8p, f32p, f32p, f32p, f32p, f32p, i64, i64, i64, ctypes.c_float]
        lib.onebit_oracle_forward_dense.argtypes = [f32p, i8p, f32p, f32p, f32p, f32p, f32p, i64, i64, i64,
                                                    ctypes.c_float]
        lib.onebit_oracle_num_threads.restype = ctypes.c_int
        for fn in (lib.onebit_oracle_pack, lib.onebit_oracle_unpack, lib.onebit_oracle_forward,
                   lib.onebit_oracle_forward_dense):
            fn.restype = None
        _lib = lib
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a, ctype):
    return a.ctypes.data_as(ctypes.POINTER(ctype)) if a is not None else None


# ----------------------------------------------------------------------------------------------
# numpy restatement
# ----------------------------------------------------------------------------------------------
def pack_signs_np(signs: np.ndarray) -> np.ndarray:
    """convert_llama_to_infer_ckpt.py:7-15. `signs` is [N, K] with entries in {+1, -1, 0}; returns int8 [N, K/8]."""
    signs = np.asarray(signs, dtype=np.float32)
    n, k = signs.shape
    assert k % 8 == 0
    bits = ((0.0 - signs + 1.0) / 2.0).astype(np.uint8)  # +1 -> 0, -1 -> 1, 0 -> 0 (0.5 truncates)
    return np.packbits(bits.reshape(n, k // 8, 8), axis=-1, bitorder="little").reshape(n, k // 8).view(np.int8)


def unpack_signs_np(packed: np.ndarray) -> np.ndarray:
    """bitnet.py:98-110. int8 [N, K/8] -> float32 [N, K] of +1 / -1 (bit 1 <=> -1, LSB first)."""
    packed = np.ascontiguousarray(packed).view(np.uint8)
    bits = np.unpackbits(packed[:, :, None], axis=-1, bitorder="little").reshape(packed.shape[0], -1)
    return (1.0 - 2.0 * bits.astype(np.float32)).astype(np.float32)


def layernorm_np(u: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    """nn.LayerNorm(N, elementwise_affine=False): biased variance over the last axis (bitnet.py:86,118)."""
    u64 = u.astype(np.float64)
    mean = u64.mean(-1, keepdims=True)
    var = ((u64 - mean) ** 2).mean(-1, keepdims=True)
    return ((u64 - mean) / np.sqrt(var + eps)).astype(np.float32)


def bitlinear_forward_np(x, packed, g, h, bias=None, eps: float = 1e-5, return_pre_ln: bool = False):
    """bitnet.py:112-122 in float32 (float64 accumulation). x [..., K] -> y [..., N]."""
    x = _f32(x)
    lead = x.shape[:-1]
    k = x.shape[-1]
    xp = (x.reshape(-1, k) * _f32(h)[None, :]).astype(np.float32)          # :113
    s = unpack_signs_np(packed)                                             # :114
    out = (xp.astype(np.float64) @ s.T.astype(np.float64)).astype(np.float32)  # :115
    u = (out * _f32(g)[None, :]).astype(np.float32)                         # :116
    y = layernorm_np(u, eps)                                                # :118
    if bias is not None:
        y = (y + _f32(bias)[None, :]).astype(np.float32)                    # :119-120
    y = y.reshape(*lead, -1)
    return (y, u.reshape(*lead, -1)) if return_pre_ln else y


# ----------------------------------------------------------------------------------------------
# C restatement (ctypes)
# ----------------------------------------------------------------------------------------------
def pack_signs_c(signs: np.ndarray) -> np.ndarray:
    signs = _f32(signs)
    n, k = signs.shape
    out = np.empty((n, k // 8), dtype=np.int8)
    _load().onebit_oracle_pack(_ptr(signs, ctypes.c_float), _ptr(out, ctypes.c_int8), n, k)
    return out


def unpack_signs_c(packed: np.ndarray) -> np.ndarray:
    packed = np.ascontiguousarray(packed, dtype=np.int8)
    n, kb = packed.shape
    out = np.empty((n, kb * 8), dtype=np.float32)
    _load().onebit_oracle_unpack(_ptr(packed, ctypes.c_int8), _ptr(out, ctypes.c_float), n, kb * 8)
    return out


def bitlinear_forward_c(x, packed, g, h, bias=None, eps: float = 1e-5, return_pre_ln: bool = False,
                        dense: bool = False):
    """C oracle. `dense=True` runs the faithful-cost variant (materialises the +-1 matrix per call, like the
    reference does) — used only as the timed CPU baseline."""
    x = _f32(x)
    lead = x.shape[:-1]
    k = x.shape[-1]
    x2 = np.ascontiguousarray(x.reshape(-1, k))
    packed = np.ascontiguousarray(packed, dtype=np.int8)
    n = packed.shape[0]
    assert packed.shape[1] * 8 == k
    g = _f32(g)
    h = _f32(h)
    b = _f32(bias) if bias is not None else None
    y = np.empty((x2.shape[0], n), dtype=np.float32)
    lib = _load()
    if dense:
        scratch = np.empty((n, k), dtype=np.float32)
        lib.onebit_oracle_forward_dense(_ptr(x2, ctypes.c_float), _ptr(packed, ctypes.c_int8), _ptr(g, ctypes.c_float),
                                        _ptr(h, ctypes.c_float), _ptr(b, ctypes.c_float), _ptr(y, ctypes.c_float),
                                        _ptr(scratch, ctypes.c_float), x2.shape[0], k, n, eps)
        return y.reshape(*lead, n)
    u = np.empty_like(y) if return_pre_ln else None
    lib.onebit_oracle_forward(_ptr(x2, ctypes.c_float), _ptr(packed, ctypes.c_int8), _ptr(g, ctypes.c_float),
                              _ptr(h, ctypes.c_float), _ptr(b, ctypes.c_float), _ptr(y, ctypes.c_float),
                              _ptr(u, ctypes.c_float), x2.shape[0], k, n, eps)
    y = y.reshape(*lead, n)
    return (y, u.reshape(*lead, n)) if return_pre_ln else y


def num_threads() -> int:
    return int(_load().onebit_oracle_num_threads())


# ----------------------------------------------------------------------------------------------
# Seeded synthetic inputs (SURVEY.md §8c/§8d recipe) — shared by golden generation, tests and bench
# ----------------------------------------------------------------------------------------------
LLAMA_SHAPES = {  # name -> (K, N)
    "7b_attn": (4096, 4096),
    "7b_gate_up": (4096, 11008),
    "7b_down": (11008, 4096),
    "13b_attn": (5120, 5120),
    "13b_gate_up": (5120, 13824),
    "13b_down": (13824, 5120),
}


def synth_case(seed: int, k: int, n: int, m: int, with_bias: bool = False):
    """weight bytes uniform over int8, g~U(0.5,1.5), h~U(-1.5,1.5) (negative h included), x~N(0,1);
    all floats rounded to fp16-representable values so fp16/bf16/fp32 kernels see identical inputs."""
    rng = np.random.Generator(np.random.PCG64(seed))
    packed = rng.integers(-128, 128, size=(n, k // 8), dtype=np.int8)
    g = rng.uniform(0.5, 1.5, size=n).astype(np.float16).astype(np.float32)
    h = rng.uniform(-1.5, 1.5, size=k).astype(np.float16).astype(np.float32)
    x = rng.standard_normal(size=(m, k)).astype(np.float16).astype(np.float32)
    bias = rng.uniform(-0.1, 0.1, size=n).astype(np.float16).astype(np.float32) if with_bias else None
    return {"x": x, "packed": packed, "g": g, "h": h, "bias": bias}


def rel_l2(a: np.ndarray, b: np.ndarray) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
