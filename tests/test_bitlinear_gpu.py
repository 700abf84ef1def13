"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the committed golden vectors.

Tolerance (BASELINE.json north_star): rel-L2 <= 1e-3 against the reference's fp32 path on identical
fp16-representable inputs; bit-exact for the integer/bit-layout work (pack / unpack).
"""
import ctypes

import numpy as np
import pytest
import torch

import onebit_b200
from onebit_b200 import BitLinearB200, _lib
from oracle import oracle

pytestmark = pytest.mark.gpu

REL_TOL = 1e-3
DTYPES = {"f16": torch.float16, "bf16": torch.bfloat16, "f32": torch.float32}
# bf16 cannot represent the fp16-rounded inputs exactly; its own rounding of y (8 mantissa bits) dominates.
OUT_TOL = {"f16": REL_TOL, "f32": REL_TOL, "bf16": 4e-3}


def dev():
    return torch.device("cuda:0")


def variants_for(m, k, n, dtype=torch.float16):
    out = ["simt"]
    lib = _lib.load()
    # probing with the forced variant: unsupported shapes return ONEBIT_ERR_INVALID_ARGUMENT before any launch
    t = torch.empty(max(m, 1) * n, dtype=torch.float32, device=dev())
    x = torch.zeros(max(m, 1) * k, dtype=dtype, device=dev())
    w = torch.zeros(n * k // 8, dtype=torch.int8, device=dev())
    g = torch.ones(max(n, k), dtype=dtype, device=dev())
    code = {torch.float16: 0, torch.bfloat16: 1, torch.float32: 2}[dtype]
    wsb = lib.onebit_matvec_workspace_bytes(max(m, 1), k)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev())
    rc = lib.onebit_bitlinear_matvec(x.data_ptr(), w.data_ptr(), g.data_ptr(), g.data_ptr(), t.data_ptr(), m, k, n,
                                     code, code, 0, ws.data_ptr(), wsb, _lib.VARIANT_MMA, None)
    torch.cuda.synchronize()
    if rc == 0:
        out.append("mma")
    rc = lib.onebit_bitlinear_matvec(x.data_ptr(), w.data_ptr(), g.data_ptr(), g.data_ptr(), t.data_ptr(), m, k, n,
                                     code, code, 0, ws.data_ptr(), wsb, _lib.VARIANT_TC5, None)
    torch.cuda.synchronize()
    if rc == 0:
        out.append("tc5")
    return out


def to_dev(case, act_dtype, param_dtype=None):
    param_dtype = param_dtype or act_dtype
    d = dev()
    return dict(
        x=torch.from_numpy(case["x"]).to(d, act_dtype),
        w=torch.from_numpy(case["packed"]).to(d),
        g=torch.from_numpy(case["g"]).to(d, param_dtype),
        h=torch.from_numpy(case["h"]).to(d, param_dtype),
        b=torch.from_numpy(case["bias"]).to(d, param_dtype) if case["bias"] is not None else None,
    )


def run(case, act="f16", variant="auto", param=None):
    t = to_dev(case, DTYPES[act], DTYPES[param] if param else None)
    y = onebit_b200.bitlinear_forward(t["x"], t["w"], t["g"], t["h"], t["b"], variant=variant)
    torch.cuda.synchronize()
    assert y.dtype == DTYPES[act] and y.device == t["x"].device
    return y.float().cpu().numpy()


def test_device_is_b200_and_library_loaded():
    assert _lib.load().onebit_device_check(0) == 0, _lib.last_error()
    assert torch.cuda.get_device_capability(0)[0] == 10


@pytest.mark.parametrize("act", ["f16", "bf16", "f32"])
def test_small_golden_cases(golden_dir, act):
    fwd = np.load(golden_dir / "forward_golden.npz")
    for seed, k, n, m, bias in fwd["small_meta"]:
        case = oracle.synth_case(int(seed), int(k), int(n), int(m), with_bias=bool(bias))
        if act == "bf16":  # feed bf16-representable inputs so only the kernel is under test
            for key in ("x", "g", "h", "bias"):
                if case[key] is not None:
                    case[key] = torch.from_numpy(case[key]).bfloat16().float().numpy()
            want = oracle.bitlinear_forward_c(case["x"], case["packed"], case["g"], case["h"], case["bias"])
        else:
            want = fwd[f"small_{seed}_y"]  # produced by the reference itself
        for variant in variants_for(int(m), int(k), int(n), DTYPES[act]):
            got = run(case, act, variant)
            assert got.shape == want.shape
            assert oracle.rel_l2(got, want) < OUT_TOL[act], (int(seed), act, variant, oracle.rel_l2(got, want))


def test_full_llama_shapes_against_reference_golden(golden_dir):
    fwd = np.load(golden_dir / "forward_golden.npz")
    worst = 0.0
    for (seed, k, n, m), name in zip(fwd["full_meta"], fwd["full_names"]):
        case = oracle.synth_case(int(seed), int(k), int(n), int(m))
        want = fwd[f"full_{name}_y"]
        for variant in variants_for(int(m), int(k), int(n)):
            got = run(case, "f16", variant)
            r = oracle.rel_l2(got, want)
            worst = max(worst, r)
            assert r < REL_TOL, (name, variant, r)
            assert np.abs(got - want).max() < 2e-2, (name, variant)
    print("worst rel-L2 over full shapes:", worst)


@pytest.mark.parametrize("m", [1, 2, 5, 8, 17, 32, 64])
def test_batch_sizes_against_oracle(m):
    for (k, n) in [(4096, 4096), (5120, 13824)]:
        case = oracle.synth_case(1000 + m, k, n, m)
        want = oracle.bitlinear_forward_c(case["x"], case["packed"], case["g"], case["h"])
        for variant in variants_for(m, k, n):
            got = run(case, "f16", variant)
            assert oracle.rel_l2(got, want) < REL_TOL, (m, k, n, variant, oracle.rel_l2(got, want))


def test_fp32_activations_are_tighter():
    case = oracle.synth_case(77, 4096, 4096, 3)
    want = oracle.bitlinear_forward_c(case["x"], case["packed"], case["g"], case["h"])
    got = run(case, "f32", "simt")
    assert oracle.rel_l2(got, want) < 2e-5


def test_mixed_param_dtype_fp32_params_fp16_activations():
    case = oracle.synth_case(78, 1024, 256, 4, with_bias=True)
    want = oracle.bitlinear_forward_c(case["x"], case["packed"], case["g"], case["h"], case["bias"])
    got = run(case, "f16", "auto", param="f32")
    assert oracle.rel_l2(got, want) < REL_TOL


def test_3d_input_and_noncontiguous_input():
    case = oracle.synth_case(21, 256, 48, 6)
    t = to_dev(case, torch.float16)
    want = oracle.bitlinear_forward_c(case["x"], case["packed"], case["g"], case["h"])
    y = onebit_b200.bitlinear_forward(t["x"].reshape(2, 3, 256), t["w"], t["g"], t["h"])
    assert tuple(y.shape) == (2, 3, 48)
    assert oracle.rel_l2(y.float().cpu().numpy().reshape(6, 48), want) < REL_TOL
    xt = t["x"].t().contiguous().t()  # same values, column-major strides
    assert not xt.is_contiguous()
    y2 = onebit_b200.bitlinear_forward(xt, t["w"], t["g"], t["h"])
    assert torch.equal(y2, y.reshape(6, 48))


def test_edge_bytes_pre_layernorm():
    k, n = 512, 40
    packed = np.zeros((n, k // 8), dtype=np.int8)
    packed[1, :] = -1        # all -1
    packed[2, :] = -128      # only bit 7 set (0x80: sign-extension hazard in the reference's >>)
    packed[3, ::2] = 0x55
    rng = np.random.Generator(np.random.PCG64(5))
    packed[4:] = rng.integers(-128, 128, size=(n - 4, k // 8), dtype=np.int8)
    x = rng.standard_normal((3, k)).astype(np.float16).astype(np.float32)
    g = rng.uniform(0.5, 1.5, n).astype(np.float16).astype(np.float32)
    h = rng.uniform(-1.5, 1.5, k).astype(np.float16).astype(np.float32)
    _, want_u = oracle.bitlinear_forward_c(x, packed, g, h, return_pre_ln=True)
    d = dev()
    for variant in variants_for(3, k, n):
        t = onebit_b200.bitlinear_matvec(torch.from_numpy(x).to(d, torch.float16), torch.from_numpy(packed).to(d),
                                         torch.from_numpy(g).to(d, torch.float16),
                                         torch.from_numpy(h).to(d, torch.float16), scale_by_g=True, variant=variant)
        assert oracle.rel_l2(t.cpu().numpy(), want_u) < 2e-4, variant


def test_odd_shapes_take_the_generic_path():
    # K = 136 (17 bytes per row: not a multiple of 4 -> byte path), N = 33, M = 7; K = 8 (one byte per row)
    for seed, k, n, m in [(7, 136, 33, 7), (6, 8, 4, 1), (9, 172 * 8, 50, 3), (10, 216 * 8, 70, 9)]:
        case = oracle.synth_case(seed, k, n, m, with_bias=(seed == 7))
        want = oracle.bitlinear_forward_c(case["x"], case["packed"], case["g"], case["h"], case["bias"])
        got = run(case, "f16", "auto")
        assert oracle.rel_l2(got, want) < REL_TOL, (k, n, m)


def test_empty_batch():
    m = BitLinearB200(64, 16, dtype=torch.float16).to(dev())
    y = m(torch.empty(0, 64, dtype=torch.float16, device=dev()))
    assert tuple(y.shape) == (0, 16)
    y = m(torch.empty(2, 0, 64, dtype=torch.float16, device=dev()))
    assert tuple(y.shape) == (2, 0, 16)


def test_pack_unpack_bit_exact(golden_dir):
    kat = np.load(golden_dir / "pack_kat.npz")
    d = dev()
    for name in ("random", "allbytes", "with_zero"):
        signs, packed = kat[f"{name}_signs"], kat[f"{name}_packed"]
        for dt in (torch.float16, torch.bfloat16, torch.float32):
            got = onebit_b200.pack_signs(torch.from_numpy(signs).to(d, dt)).cpu().numpy()
            np.testing.assert_array_equal(got, packed)
            un = onebit_b200.unpack_signs(torch.from_numpy(packed).to(d), dt).float().cpu().numpy()
            np.testing.assert_array_equal(un, kat[f"{name}_unpacked"])
    lat = torch.from_numpy(kat["latent_fp16"]).to(d)
    np.testing.assert_array_equal(onebit_b200.pack_signs(torch.sign(lat)).cpu().numpy(), kat["latent_packed"])
    # round trip at a full LLaMA shape: pack(unpack(w)) == w
    w = torch.randint(-128, 128, (4096, 11008 // 8), dtype=torch.int8, device=d)
    assert torch.equal(onebit_b200.pack_signs(onebit_b200.unpack_signs(w, torch.float16)), w)


def test_train_inf_equivalence_fixture(golden_dir):
    z = np.load(golden_dir / "train_inf_equiv.npz")
    d = dev()
    packed = onebit_b200.pack_signs(torch.sign(torch.from_numpy(z["latent"]).to(d)))
    np.testing.assert_array_equal(packed.cpu().numpy(), z["packed"])
    y = onebit_b200.bitlinear_forward(torch.from_numpy(z["x"]).to(d), packed, torch.from_numpy(z["g"]).to(d),
                                      torch.from_numpy(z["h"]).to(d))
    assert oracle.rel_l2(y.cpu().numpy(), z["y_inf"]) < 2e-5


def test_linearity_and_sign_flip_properties_at_full_size():
    # size-independent properties at BASELINE sizes (no oracle needed): t is linear in x; flipping every
    # weight bit negates t; LayerNorm output has zero mean / unit variance per token.
    d = dev()
    k, n, m = 13824, 5120, 4
    gen = torch.Generator(device="cpu").manual_seed(3)
    w = torch.randint(-128, 128, (n, k // 8), dtype=torch.int8, generator=gen).to(d)
    g = (torch.rand(n, generator=gen) + 0.5).half().to(d)
    h = (torch.rand(k, generator=gen) * 3 - 1.5).half().to(d)
    x1 = torch.randn(m, k, generator=gen).half().to(d)
    x2 = torch.randn(m, k, generator=gen).half().to(d)
    for variant in variants_for(m, k, n):
        t1 = onebit_b200.bitlinear_matvec(x1, w, g, h, variant=variant)
        t2 = onebit_b200.bitlinear_matvec(x2, w, g, h, variant=variant)
        t12 = onebit_b200.bitlinear_matvec((x1.float() + x2.float()).half(), w, g, h, variant=variant)
        scale = t1.abs().mean().item()
        assert (t12 - (t1 + t2)).abs().max().item() < 0.05 * scale
        tneg = onebit_b200.bitlinear_matvec(x1, ~w, g, h, variant=variant)
        assert torch.allclose(tneg, -t1, rtol=0, atol=1e-3 * scale)
        y = onebit_b200.bitlinear_forward(x1, w, g, h, variant=variant).float()
        assert y.mean(-1).abs().max().item() < 2e-3
        assert (y.var(-1, unbiased=False) - 1).abs().max().item() < 5e-3


def test_module_forward_matches_functional_and_reference_semantics():
    case = oracle.synth_case(31, 1024, 320, 5, with_bias=True)
    want = oracle.bitlinear_forward_c(case["x"], case["packed"], case["g"], case["h"], case["bias"])
    mod = BitLinearB200(1024, 320, bias=True, dtype=torch.float16)
    with torch.no_grad():
        mod.weight.copy_(torch.from_numpy(case["packed"]))
        mod.weight_scale.copy_(torch.from_numpy(case["g"]))
        mod.input_factor.copy_(torch.from_numpy(case["h"]))
        mod.bias.copy_(torch.from_numpy(case["bias"]))
    mod = mod.to(dev())
    assert mod.weight.dtype == torch.int8
    y = mod(torch.from_numpy(case["x"]).to(dev(), torch.float16))
    assert oracle.rel_l2(y.float().cpu().numpy(), want) < REL_TOL
    # .float() module with float input, like the reference's CPU fp32 use
    y32 = mod.float()(torch.from_numpy(case["x"]).to(dev()))
    assert y32.dtype == torch.float32
    assert oracle.rel_l2(y32.cpu().numpy(), want) < REL_TOL


def test_error_behaviour_on_gpu():
    d = dev()
    mod = BitLinearB200(64, 16, dtype=torch.float16).to(d)
    with pytest.raises(RuntimeError, match="features"):
        mod(torch.zeros(2, 32, dtype=torch.float16, device=d))
    with pytest.raises(RuntimeError, match="dtype"):
        mod(torch.zeros(2, 64, dtype=torch.float64, device=d))
    with pytest.raises(RuntimeError, match="int8"):
        onebit_b200.bitlinear_forward(torch.zeros(2, 64, dtype=torch.float16, device=d),
                                      torch.zeros(16, 8, dtype=torch.uint8, device=d), mod.weight_scale,
                                      mod.input_factor)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mod(torch.zeros(2, 64, dtype=torch.float16))


def test_cuda_graph_capture_and_replay():
    d = dev()
    case = oracle.synth_case(41, 4096, 4096, 2)
    t = to_dev(case, torch.float16)
    want = oracle.bitlinear_forward_c(case["x"], case["packed"], case["g"], case["h"])
    static_x = t["x"].clone()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            onebit_b200.bitlinear_forward(static_x, t["w"], t["g"], t["h"])
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        static_y = onebit_b200.bitlinear_forward(static_x, t["w"], t["g"], t["h"])
    static_x.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.isnan(static_y).all() or static_y.abs().max() < 1e-3  # LN of an all-zero row: 0/sqrt(eps)
    static_x.copy_(t["x"])
    graph.replay()
    torch.cuda.synchronize()
    assert oracle.rel_l2(static_y.float().cpu().numpy(), want) < REL_TOL
    assert d.type == "cuda"


def test_concurrent_streams_share_weights():
    case = oracle.synth_case(42, 4096, 4096, 1)
    t = to_dev(case, torch.float16)
    want = oracle.bitlinear_forward_c(case["x"], case["packed"], case["g"], case["h"])
    streams = [torch.cuda.Stream() for _ in range(4)]
    outs = []
    torch.cuda.synchronize()
    for s in streams:
        with torch.cuda.stream(s):
            for _ in range(8):
                y = onebit_b200.bitlinear_forward(t["x"], t["w"], t["g"], t["h"])
            outs.append(y)
    torch.cuda.synchronize()
    for y in outs:
        assert oracle.rel_l2(y.float().cpu().numpy(), want) < REL_TOL
    assert all(torch.equal(outs[0], y) for y in outs[1:])  # deterministic


def test_host_buffer_layer_handle():
    lib = _lib.load()
    case = oracle.synth_case(43, 4096, 4096, 3)
    want = oracle.bitlinear_forward_c(case["x"], case["packed"], case["g"], case["h"])
    g16, h16 = case["g"].astype(np.float16), case["h"].astype(np.float16)
    x16 = torch.from_numpy(case["x"].astype(np.float16)).pin_memory()
    y16 = torch.empty((3, 4096), dtype=torch.float16).pin_memory()
    handle = ctypes.c_void_p()
    rc = lib.onebit_layer_create(ctypes.byref(handle), case["packed"].ctypes.data, g16.ctypes.data, h16.ctypes.data, None,
                                 4096, 4096, _lib.F16, _lib.F16, 1e-5, 8)
    assert rc == 0, _lib.last_error()
    try:
        rc = lib.onebit_layer_forward_host(handle, x16.data_ptr(), y16.data_ptr(), 3, None)
        assert rc == 0, _lib.last_error()
        assert oracle.rel_l2(y16.float().numpy(), want) < REL_TOL
        assert lib.onebit_layer_forward_host(handle, x16.data_ptr(), y16.data_ptr(), 9, None) == -1  # > max_m
    finally:
        lib.onebit_layer_destroy(handle)


def test_tensor_parallel_halves_compose_to_the_full_layer():
    # column-parallel (N shards): partial (sum, sumsq) -> sum over shards -> apply; row-parallel (K shards):
    # partial t summed over shards -> scale + LayerNorm. Single GPU stands in for the ranks here; the
    # gloo test in test_tp_cpu.py covers the collective plumbing.
    lib = _lib.load()
    d = dev()
    case = oracle.synth_case(44, 1024, 512, 4)
    want = oracle.bitlinear_forward_c(case["x"], case["packed"], case["g"], case["h"])
    t = to_dev(case, torch.float16)
    # N shards
    shards = 4
    ns = 512 // shards
    stats = torch.zeros(4, 2, dtype=torch.float64, device=d)
    ts = []
    for r in range(shards):
        tr = onebit_b200.bitlinear_matvec(t["x"], t["w"][r * ns:(r + 1) * ns].contiguous(),
                                          t["g"][r * ns:(r + 1) * ns].contiguous(), t["h"], scale_by_g=False)
        st = torch.empty(4, 2, dtype=torch.float64, device=d)
        gr = t["g"][r * ns:(r + 1) * ns].contiguous()
        assert lib.onebit_scale_partial_stats(tr.data_ptr(), gr.data_ptr(), st.data_ptr(), 4, ns, _lib.F16, None) == 0
        stats += st
        ts.append((tr, gr))
    outs = []
    for tr, gr in ts:
        y = torch.empty(4, ns, dtype=torch.float16, device=d)
        assert lib.onebit_layernorm_apply_stats(tr.data_ptr(), gr.data_ptr(), None, stats.data_ptr(), y.data_ptr(), 4,
                                                ns, 512, _lib.F16, _lib.F16, 1e-5, None) == 0
        outs.append(y)
    got = torch.cat(outs, dim=1).float().cpu().numpy()
    assert oracle.rel_l2(got, want) < REL_TOL
    # K shards
    ks = 1024 // shards
    tsum = torch.zeros(4, 512, dtype=torch.float32, device=d)
    for r in range(shards):
        wr = t["w"][:, r * ks // 8:(r + 1) * ks // 8].contiguous()
        tsum += onebit_b200.bitlinear_matvec(t["x"][:, r * ks:(r + 1) * ks].contiguous(), wr, t["g"],
                                             t["h"][r * ks:(r + 1) * ks].contiguous(), scale_by_g=False)
    y = onebit_b200.scale_layernorm(tsum, t["g"], None, torch.float16)
    assert oracle.rel_l2(y.float().cpu().numpy(), want) < REL_TOL


@pytest.mark.parametrize("m", [9, 16, 100, 256, 257, 600])
def test_prefill_tcgen05_variant_against_oracle(m):
    # tcgen05 / TMEM path: token tails (m % 256, m % 16), weight-row tails (N % 256), several K
    for (k, n) in [(4096, 4096), (11008, 4096), (1024, 300), (64, 256)]:
        if m > 300 and k * n > 1024 * 1024:
            continue  # keep the CPU oracle to a few seconds
        case = oracle.synth_case(2000 + m, k, n, m)
        want = oracle.bitlinear_forward_c(case["x"], case["packed"], case["g"], case["h"])
        got = run(case, "f16", "tc5")
        r = oracle.rel_l2(got, want)
        assert r < REL_TOL, (m, k, n, r)


def test_prefill_tcgen05_pre_layernorm_and_dtypes():
    case = oracle.synth_case(2100, 2048, 512, 130, with_bias=True)
    _, want_u = oracle.bitlinear_forward_c(case["x"], case["packed"], case["g"], case["h"], case["bias"], return_pre_ln=True)
    d = dev()
    t = onebit_b200.bitlinear_matvec(torch.from_numpy(case["x"]).to(d, torch.float16), torch.from_numpy(case["packed"]).to(d),
                                     torch.from_numpy(case["g"]).to(d, torch.float16),
                                     torch.from_numpy(case["h"]).to(d, torch.float16), scale_by_g=True, variant="tc5")
    assert oracle.rel_l2(t.cpu().numpy(), want_u) < 1e-5  # fp16 x fp16 products are exact in the fp32 accumulator
    want = oracle.bitlinear_forward_c(case["x"], case["packed"], case["g"], case["h"], case["bias"])
    for act in ("bf16", "f32"):
        c2 = dict(case)
        if act == "bf16":
            for key in ("x", "g", "h", "bias"):
                c2[key] = torch.from_numpy(case[key]).bfloat16().float().numpy()
            w2 = oracle.bitlinear_forward_c(c2["x"], c2["packed"], c2["g"], c2["h"], c2["bias"])
        else:
            w2 = want
        got = run(c2, act, "tc5")
        assert oracle.rel_l2(got, w2) < OUT_TOL[act], act


def test_auto_dispatch_picks_a_tensor_core_variant_for_every_llama_batch():
    # AUTO must agree with the forced variants (decode: IMMA, prefill: tcgen05) on the same inputs
    case = oracle.synth_case(2200, 4096, 4096, 40)
    a = run(case, "f16", "auto")
    b = run(case, "f16", "tc5")
    assert np.array_equal(a, b)
    case = oracle.synth_case(2201, 4096, 4096, 2)
    assert np.array_equal(run(case, "f16", "auto"), run(case, "f16", "mma"))


def test_tcgen05_bfloat16_activations_keep_their_range():
    """bf16 activations used to be rounded to fp16 on the tcgen05 path (|x| > 65504 -> inf, tiny values flushed).
    They now stay bf16 through the MMA (kind::f16 with BF16 operands): inputs far outside the fp16 range must match the
    oracle evaluated on the same bf16 values."""
    case = oracle.synth_case(2300, 1024, 384, 40)
    x = case["x"].copy()
    x[::3] *= 3.0e5     # beyond the fp16 maximum
    x[1::3] *= 1.0e-7   # below the fp16 subnormal range
    c2 = dict(case, x=torch.from_numpy(x).bfloat16().float().numpy())
    for key in ("g", "h"):
        c2[key] = torch.from_numpy(case[key]).bfloat16().float().numpy()
    d = dev()
    t = onebit_b200.bitlinear_matvec(torch.from_numpy(c2["x"]).to(d, torch.bfloat16), torch.from_numpy(c2["packed"]).to(d),
                                     torch.from_numpy(c2["g"]).to(d, torch.bfloat16), torch.from_numpy(c2["h"]).to(d, torch.bfloat16),
                                     scale_by_g=True, variant="tc5")
    _, want_u = oracle.bitlinear_forward_c(c2["x"], c2["packed"], c2["g"], c2["h"], None, return_pre_ln=True)
    got = t.cpu().numpy()
    assert np.isfinite(got).all()
    for rows in (slice(0, None, 3), slice(1, None, 3), slice(2, None, 3)):  # per magnitude class (a global norm would hide the small rows)
        assert oracle.rel_l2(got[rows], want_u[rows]) < 1e-5
    # fp16 parameters next to bf16 activations: input_factor is re-rounded to bf16 (documented), result stays finite and close
    y = run(dict(c2, bias=None), "bf16", "tc5", param="f16")
    assert np.isfinite(y).all()


def test_auto_dispatch_for_small_batches_at_wide_k_matches_the_oracle():
    """AUTO must not land on the CUDA-core anchor for 5..8 tokens at widths whose digits do not fit the IMMA GEMV's shared
    memory (K = 11008: down_proj): those go to the tcgen05 tile. Checked through the public forward."""
    for m in (5, 8):
        k, n = 11008, 4096
        assert "mma" not in variants_for(m, k, n)  # the reason the case exists
        case = oracle.synth_case(4200 + m, k, n, m)
        want = oracle.bitlinear_forward_c(case["x"], case["packed"], case["g"], case["h"])
        got_auto, got_tc5 = run(case, "f16", "auto"), run(case, "f16", "tc5")
        assert oracle.rel_l2(got_auto, want) < REL_TOL
        assert np.array_equal(got_auto, got_tc5)  # AUTO picked the tcgen05 variant


def test_infinite_input_poisons_the_token_instead_of_finite_garbage():
    """An Inf in a token's input makes that token's outputs non-finite (the reference's fp arithmetic propagates it);
    the other token of the batch is untouched (decode GEMV quantiser)."""
    k, n, m = 4096, 4096, 2
    case = oracle.synth_case(77, k, n, m)
    want = oracle.bitlinear_forward_c(case["x"], case["packed"], case["g"], case["h"])
    t = to_dev(case, torch.float16)
    t["x"][1, 5] = float("inf")
    y = onebit_b200.bitlinear_forward(t["x"], t["w"], t["g"], t["h"], t["b"], variant="mma").float().cpu().numpy()
    assert not np.isfinite(y[1]).any()
    assert oracle.rel_l2(y[0], want[0]) < REL_TOL
