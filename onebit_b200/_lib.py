"""ctypes binding of libonebit_b200.so (the C ABI in include/onebit_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
from pathlib import Path

PKG = Path(__file__).resolve().parent
import os as _os

LIB_PATH = PKG / f"libonebit_b200{_os.environ.get('ONEBIT_LIB_SUFFIX', '')}.so"

OK = 0
F16, BF16, F32 = 0, 1, 2
VARIANT_AUTO, VARIANT_SIMT, VARIANT_MMA, VARIANT_TC5 = 0, 1, 2, 3
VARIANTS = {"auto": VARIANT_AUTO, "simt": VARIANT_SIMT, "mma": VARIANT_MMA, "tc5": VARIANT_TC5}

_c = ctypes
_vp, _i64, _int, _f32, _sz = _c.c_void_p, _c.c_int64, _c.c_int, _c.c_float, _c.c_size_t

ALLREDUCE_FN = _c.CFUNCTYPE(_int, _vp, _vp, _i64, _vp)  # int (*)(void* user, float* data, int64 count, void* stream)

# name -> (restype, argtypes); must list every symbol include/onebit_b200.h declares (checked by tests)
SIGNATURES = {
    "onebit_version": (_c.c_char_p, []),
    "onebit_last_error": (_c.c_char_p, []),
    "onebit_device_check": (_int, [_int]),
    "onebit_pack_signs": (_int, [_vp, _vp, _i64, _i64, _int, _vp]),
    "onebit_unpack_signs": (_int, [_vp, _vp, _i64, _i64, _int, _vp]),
    "onebit_bitlinear_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "onebit_bitlinear_forward": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _int, _int, _f32, _vp, _sz,
                                        _int, _vp]),
    "onebit_matvec_workspace_bytes": (_sz, [_i64, _i64]),
    "onebit_bitlinear_matvec": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _int, _int, _int, _vp, _sz, _int,
                                       _vp]),
    "onebit_scale_layernorm": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _int, _int, _f32, _vp]),
    "onebit_scale_partial_stats": (_int, [_vp, _vp, _vp, _i64, _i64, _int, _vp]),
    "onebit_layernorm_apply_stats": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _int, _int, _f32, _vp]),
    "onebit_layer_create": (_int, [_c.POINTER(_vp), _vp, _vp, _vp, _vp, _i64, _i64, _int, _int, _f32, _i64]),
    "onebit_layer_forward_host": (_int, [_vp, _vp, _vp, _i64, _vp]),
    "onebit_layer_forward_device": (_int, [_vp, _vp, _vp, _i64, _vp]),
    "onebit_layer_destroy": (None, [_vp]),
    "onebit_decoder_create": (_int, [_c.POINTER(_vp), _vp, _vp, _vp, _vp, _vp, _vp, _vp, ALLREDUCE_FN, _vp]),
    "onebit_decoder_reset": (_int, [_vp, _vp, _vp, _int, _vp]),
    "onebit_decoder_step": (_int, [_vp, _int, _vp, _vp, _vp]),
    "onebit_decoder_step_host": (_int, [_vp, _int, _vp, _vp, _vp]),
    "onebit_decoder_gemv_only": (_int, [_vp, _int, _vp]),
    "onebit_decoder_next_ids": (_vp, [_vp]),
    "onebit_decoder_positions": (_vp, [_vp]),
    "onebit_decoder_kernel_launches_per_step": (_int, [_vp]),
    "onebit_decoder_is_persistent": (_int, [_vp]),
    "onebit_decoder_prefill": (_int, [_vp, _int, _int, _int, _vp, _vp, _vp, _vp]),
    "onebit_decoder_enable_p2p_allreduce": (_int, [_vp, _int, _int, _vp, _sz]),
    "onebit_decoder_status": (_int, [_vp, _c.POINTER(_int)]),
    "onebit_decoder_read_trace": (_int, [_vp, _vp, _int]),
    "onebit_decoder_destroy": (None, [_vp]),
}


class BitLinearParams(_c.Structure):
    _fields_ = [("weight", _vp), ("weight_scale", _vp), ("input_factor", _vp)]


class LayerParams(_c.Structure):
    _fields_ = [(n, BitLinearParams) for n in ("q", "k", "v", "o", "gate", "up", "down")] + [
        ("input_layernorm", _vp), ("post_attention_layernorm", _vp)]


class DecoderConfig(_c.Structure):
    _fields_ = [(n, _int) for n in ("hidden_size", "intermediate_size", "num_layers", "num_heads", "vocab_size",
                                    "max_seq_len", "max_batch", "param_dtype")] + [
        ("rms_eps", _f32), ("ln_eps", _f32), ("tp_size", _int), ("tp_rank", _int)]

_lib = None


def load() -> ctypes.CDLL:
    """Load the shared library (once). Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m onebit_b200.build` (nvcc, sm_100a). "
                "onebit_b200 has no CPU or PyTorch fallback for the 1-bit linear path.")
        lib = ctypes.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here means header and library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


_torch_ops = None


def torch_ops():
    """`torch.ops.onebit_b200` (bitlinear, bitlinear_nolayernorm, pack_signs, unpack_signs): the dispatcher face of the
    same C ABI (csrc_torch/torch_ops.cpp). Raises RuntimeError if the shim has not been built."""
    global _torch_ops
    if _torch_ops is None:
        import torch
        load()
        path = LIB_PATH.with_name("libonebit_b200_torch.so")
        if not path.exists():
            raise RuntimeError(f"{path} is missing: build it with `python -m onebit_b200.build`")
        torch.ops.load_library(str(path))
        _torch_ops = torch.ops.onebit_b200
    return _torch_ops


def last_error() -> str:
    return load().onebit_last_error().decode()


def check(rc: int, what: str) -> None:
    if rc != OK:
        raise RuntimeError(f"{what} failed (code {rc}): {last_error()}")
