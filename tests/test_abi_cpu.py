"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol the header declares, the
nn.Module mirrors the reference's BitLinearInf surface, and the product refuses to run without CUDA."""
import re
from pathlib import Path

import pytest
import torch
from torch import nn

import onebit_b200
from onebit_b200 import BitLinearB200, _lib, replace_bitlinear

ROOT = Path(__file__).resolve().parent.parent


def header_symbols():
    text = (ROOT / "include" / "onebit_b200.h").read_text()
    return sorted(set(re.findall(r"ONEBIT_API\s+[\w\s\*]+?\b(onebit_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/onebit_b200.h but not exported"
    assert set(syms) == set(_lib.SIGNATURES), "ctypes signature table and header disagree"
    assert b"sm_100a" in lib.onebit_version()


def test_argument_errors_are_reported_without_a_gpu():
    lib = _lib.load()
    # K not a multiple of 8 -> ONEBIT_ERR_INVALID_ARGUMENT before anything touches a device
    rc = lib.onebit_bitlinear_forward(16, 16, 16, 16, None, 16, 1, 12, 4, 0, 0, 1e-5, 16, 1 << 20, 0, None)
    assert rc == -1 and "multiple of 8" in _lib.last_error()
    rc = lib.onebit_bitlinear_forward(16, 16, 16, 16, None, 16, 1, 16, 4, 7, 0, 1e-5, 16, 1 << 20, 0, None)
    assert rc == -1 and "dtype" in _lib.last_error()
    rc = lib.onebit_bitlinear_forward(16, 16, 16, 16, None, 16, 1, 16, 4, 0, 0, 1e-5, None, 0, 0, None)
    assert rc == -4
    rc = lib.onebit_bitlinear_forward(8, 16, 16, 16, None, 16, 1, 16, 4, 0, 0, 1e-5, 16, 1 << 20, 0, None)
    assert rc == -1 and "aligned" in _lib.last_error()
    assert lib.onebit_bitlinear_workspace_bytes(4, 4096, 4096) >= 4 * 4096 * 4


def test_module_surface_matches_reference_bitlinearinf():
    m = BitLinearB200(64, 24, bias=True, dtype=torch.float32)
    sd = m.state_dict()
    assert list(sd.keys()) == ["weight", "weight_scale", "input_factor", "bias"]  # bitnet.py:78-84
    assert sd["weight"].dtype == torch.int8 and tuple(sd["weight"].shape) == (24, 8)
    assert tuple(sd["weight_scale"].shape) == (24,) and tuple(sd["input_factor"].shape) == (64,)
    assert not any(p.requires_grad for p in m.parameters())
    assert (m.weight == 0).all() and (m.weight_scale == 1).all() and (m.input_factor == 1).all()
    assert m.in_features == 64 and m.out_features == 24
    nb = BitLinearB200(64, 24)
    assert nb.bias is None and list(nb.state_dict().keys()) == ["weight", "weight_scale", "input_factor"]
    h = BitLinearB200(64, 24).half()
    assert h.weight.dtype == torch.int8 and h.weight_scale.dtype == torch.float16  # int8 survives .half()
    with pytest.raises(ValueError):
        BitLinearB200(60, 8)
    # bitnet.py:86: the reference keeps an affine-free nn.LayerNorm attribute; so does the mirror (no parameters, no keys)
    assert isinstance(m.layernorm, torch.nn.LayerNorm) and not m.layernorm.elementwise_affine
    assert tuple(m.layernorm.normalized_shape) == (24,) and m.eps == m.layernorm.eps == 1e-5
    assert [n for n, _ in m.named_children()] == ["layernorm"]


def test_state_dict_round_trip():
    a = BitLinearB200(128, 16, dtype=torch.float16)
    with torch.no_grad():
        a.weight.copy_(torch.randint(-128, 128, a.weight.shape, dtype=torch.int8))
        a.weight_scale.uniform_(0.5, 1.5)
        a.input_factor.uniform_(-1.5, 1.5)
    b = BitLinearB200(128, 16, dtype=torch.float16)
    b.load_state_dict(a.state_dict())
    for k in a.state_dict():
        assert torch.equal(a.state_dict()[k], b.state_dict()[k])


def test_cpu_tensor_is_refused_loudly():
    m = BitLinearB200(64, 8, dtype=torch.float32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(2, 64))
    with pytest.raises(RuntimeError, match="CUDA"):
        onebit_b200.pack_signs(torch.ones(4, 8))


class BitLinearInf(nn.Module):
    """Stand-in with the reference's class name and parameters (the real one is not on the GPU box)."""

    def __init__(self, k, n):
        super().__init__()
        self.in_features, self.out_features, self.groups = k, n, 1
        self.weight = nn.Parameter(torch.zeros(n, k // 8, dtype=torch.int8), requires_grad=False)
        self.weight_scale = nn.Parameter(torch.ones(n), requires_grad=False)
        self.input_factor = nn.Parameter(torch.ones(k), requires_grad=False)
        self.register_parameter("bias", None)
        self.layernorm = nn.LayerNorm(n, elementwise_affine=False)


def test_replace_bitlinear_shares_parameters_and_keys():
    class Block(nn.Module):
        def __init__(self):
            super().__init__()
            self.q_proj = BitLinearInf(64, 64)
            self.mlp = nn.Sequential(BitLinearInf(64, 128), nn.SiLU(), BitLinearInf(128, 64))
            self.head = nn.Linear(64, 10)

    model = Block()
    keys_before = list(model.state_dict().keys())
    q_weight = model.q_proj.weight
    assert replace_bitlinear(model) == 3
    assert isinstance(model.q_proj, BitLinearB200) and isinstance(model.mlp[0], BitLinearB200)
    assert isinstance(model.head, nn.Linear)
    assert model.q_proj.weight is q_weight  # same Parameter object, not a copy
    assert list(model.state_dict().keys()) == keys_before
    assert replace_bitlinear(model) == 0  # idempotent
