"""Pins the CPU oracle (numpy + C restatements) against outputs of the reference's own code.

The fixtures under tests/golden/ were produced by tests/golden/gen_golden.py, which executes
xuyuzhuang11/OneBit's BitLinearInf / BitLinear / fp16_to_int8 in the build container.
"""
import numpy as np
import pytest

from oracle import oracle


@pytest.fixture(scope="module")
def kat(golden_dir):
    return np.load(golden_dir / "pack_kat.npz")


@pytest.fixture(scope="module")
def fwd(golden_dir):
    return np.load(golden_dir / "forward_golden.npz")


@pytest.mark.parametrize("name", ["random", "allbytes", "with_zero"])
def test_pack_matches_reference_packer(kat, name):
    signs, want = kat[f"{name}_signs"], kat[f"{name}_packed"]
    assert want.dtype == np.int8
    np.testing.assert_array_equal(oracle.pack_signs_np(signs), want)
    np.testing.assert_array_equal(oracle.pack_signs_c(signs), want)


@pytest.mark.parametrize("name", ["random", "allbytes", "with_zero"])
def test_unpack_matches_reference_unpack(kat, name):
    packed, want = kat[f"{name}_packed"], kat[f"{name}_unpacked"]
    np.testing.assert_array_equal(oracle.unpack_signs_np(packed), want)
    np.testing.assert_array_equal(oracle.unpack_signs_c(packed), want)


def test_pack_of_fp16_latent_sign(kat):
    lat = kat["latent_fp16"].astype(np.float32)
    np.testing.assert_array_equal(oracle.pack_signs_np(np.sign(lat)), kat["latent_packed"])
    # sign(0) packs as +1 (bit 0): first four columns of row 0 were zeroed
    assert (kat["latent_packed"].view(np.uint8)[0, 0] & 0x0F) == 0


def test_allbytes_cover_0x80_and_extremes(kat):
    p = kat["allbytes_packed"].view(np.uint8).reshape(-1)
    assert set(p.tolist()) == set(range(256))
    assert kat["allbytes_packed"].min() == -128  # 0x80 stored as negative int8


def test_bit_polarity_and_order():
    # column 8j+i <-> bit i (LSB first) of byte j, bit 1 <=> sign -1
    s = np.ones((1, 16), dtype=np.float32)
    s[0, 0] = -1
    s[0, 9] = -1
    p = oracle.pack_signs_np(s).view(np.uint8)
    assert p.tolist() == [[1, 2]]
    assert oracle.unpack_signs_np(np.array([[-128]], dtype=np.int8)).tolist() == [[1, 1, 1, 1, 1, 1, 1, -1]]


def test_small_forward_cases(fwd):
    for seed, k, n, m, bias in fwd["small_meta"]:
        case = oracle.synth_case(int(seed), int(k), int(n), int(m), with_bias=bool(bias))
        want = fwd[f"small_{seed}_y"]
        for impl in (oracle.bitlinear_forward_np, oracle.bitlinear_forward_c):
            got = impl(case["x"], case["packed"], case["g"], case["h"], case["bias"])
            assert got.shape == want.shape
            assert oracle.rel_l2(got, want) < 2e-6, (seed, impl.__name__)
            assert np.abs(got - want).max() < 2e-5
        dense = oracle.bitlinear_forward_c(case["x"], case["packed"], case["g"], case["h"], case["bias"], dense=True)
        assert oracle.rel_l2(dense, want) < 2e-6


def test_3d_input(fwd):
    case = oracle.synth_case(21, 256, 48, 6)
    got = oracle.bitlinear_forward_c(case["x"].reshape(2, 3, 256), case["packed"], case["g"], case["h"])
    assert got.shape == (2, 3, 48)
    assert oracle.rel_l2(got, fwd["x3d_21_y"]) < 2e-6


def test_full_llama_shapes(fwd):
    for (seed, k, n, m), name in zip(fwd["full_meta"], fwd["full_names"]):
        case = oracle.synth_case(int(seed), int(k), int(n), int(m))
        got = oracle.bitlinear_forward_c(case["x"], case["packed"], case["g"], case["h"])
        want = fwd[f"full_{name}_y"]
        assert oracle.rel_l2(got, want) < 5e-6, name
        assert np.abs(got - want).max() < 1e-4, name


def test_reference_fp16_path_distance_is_documented(fwd):
    # the reference's own fp16 path sits ~5e-4 rel-L2 from its fp32 path; our tolerance is 1e-3.
    d = oracle.rel_l2(fwd["fp16_7b_attn_m3_y_fp16path"], fwd["fp16_7b_attn_m3_y_fp32path"])
    assert 1e-5 < d < 1e-3


def test_train_vs_inf_equivalence(golden_dir):
    z = np.load(golden_dir / "train_inf_equiv.npz")
    np.testing.assert_array_equal(z["y_train"], z["y_inf"])
    np.testing.assert_array_equal(oracle.pack_signs_np(np.sign(z["latent"])), z["packed"])
    got = oracle.bitlinear_forward_c(z["x"], z["packed"], z["g"], z["h"])
    assert oracle.rel_l2(got, z["y_inf"]) < 2e-6


def test_edge_bytes_forward():
    # all-zero bytes (all +1), all-ones bytes (all -1), 0x80 only
    k, n = 64, 3
    packed = np.zeros((n, k // 8), dtype=np.int8)
    packed[1, :] = -1
    packed[2, :] = -128
    rng = np.random.Generator(np.random.PCG64(5))
    x = rng.standard_normal((2, k)).astype(np.float32)
    g = np.array([1.0, 2.0, 0.5], dtype=np.float32)
    h = np.ones(k, dtype=np.float32)
    _, u = oracle.bitlinear_forward_c(x, packed, g, h, return_pre_ln=True)
    tot = x.sum(-1)
    np.testing.assert_allclose(u[:, 0], tot * 1.0, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(u[:, 1], -tot * 2.0, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(u[:, 2], (tot - 2 * x[:, 7::8].sum(-1)) * 0.5, rtol=1e-5, atol=1e-5)
