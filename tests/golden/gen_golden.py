"""Generate golden vectors by EXECUTING the reference's own code (xuyuzhuang11/OneBit @ 42d6d7b).

Runs only in the build container, where /root/reference is mounted read-only; the GPU box never sees the
reference, so the outputs are committed as small .npz fixtures next to this script:

    python tests/golden/gen_golden.py            # rewrites tests/golden/*.npz

What is executed (nothing is copied into this repo):
  * `BitLinearInf` / `BitLinear` from transformers/src/transformers/models/bitnet.py, loaded by file path
    (the file imports only torch);
  * `fp16_to_int8` from scripts/convert_llama_to_infer_ckpt.py — that script loads checkpoints at import
    time, so only the function's own AST node is compiled and executed;
  * (model-level fixture) `BitLlamaForCausalLMInf` from the reference's vendored transformers, imported with
    a stub for `transformers.dependency_versions_check` (it only enforces a tokenizers version pin).

Inputs come from oracle.synth_case (numpy PCG64 seeds), so fixtures store seeds + outputs, not weights,
for the full LLaMA shapes.
"""
from __future__ import annotations

import ast
import importlib.util
import os
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = Path(os.environ.get("ONEBIT_REFERENCE", "/root/reference"))
HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.dont_write_bytecode = True

from oracle import oracle  # noqa: E402


def load_bitnet():
    path = REF / "transformers/src/transformers/models/bitnet.py"
    spec = importlib.util.spec_from_file_location("ref_bitnet", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_packer():
    path = REF / "scripts/convert_llama_to_infer_ckpt.py"
    tree = ast.parse(path.read_text())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "fp16_to_int8")
    ns = {"torch": torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), str(path), "exec"), ns)
    return ns["fp16_to_int8"]


def ref_forward(bitnet, case, dtype=torch.float32):
    k = case["x"].shape[-1]
    n = case["packed"].shape[0]
    mod = bitnet.BitLinearInf(k, n, bias=case["bias"] is not None, dtype=dtype)
    with torch.no_grad():
        mod.weight.copy_(torch.from_numpy(case["packed"]))
        mod.weight_scale.copy_(torch.from_numpy(case["g"]))
        mod.input_factor.copy_(torch.from_numpy(case["h"]))
        if case["bias"] is not None:
            mod.bias.copy_(torch.from_numpy(case["bias"]))
        y = mod(torch.from_numpy(case["x"]).to(dtype))
    return y.float().numpy()


def gen_pack(bitnet, packer):
    rng = np.random.Generator(np.random.PCG64(1234))
    out = {}
    # random +-1, plus rows that exercise every byte value, 0x80 (sign-extension in >>), sign(0) -> +1
    signs = rng.choice(np.array([-1.0, 1.0], dtype=np.float32), size=(40, 256))
    allbytes = np.arange(256, dtype=np.uint8)
    bits = np.unpackbits(allbytes[:, None], axis=1, bitorder="little")  # [256, 8]
    signs_all = (1.0 - 2.0 * bits.astype(np.float32)).reshape(8, 256)    # 32 bytes per row
    with_zero = signs.copy()
    with_zero[::3, ::5] = 0.0
    for name, s in (("random", signs), ("allbytes", signs_all), ("with_zero", with_zero)):
        packed = packer(torch.from_numpy(s)).numpy()
        out[f"{name}_signs"] = s
        out[f"{name}_packed"] = packed
        # reference unpack of what the reference packed (bitnet.py:98-110), fp32
        mod = bitnet.BitLinearInf(s.shape[1], s.shape[0], dtype=torch.float32)
        out[f"{name}_unpacked"] = mod.int8_to_fp16(torch.from_numpy(packed)).numpy()
    # fp16 input to the packer, as the convert script feeds it (sign of fp16 latent weights)
    lat = torch.from_numpy(rng.standard_normal((16, 128)).astype(np.float16))
    lat[0, :4] = 0.0
    out["latent_fp16"] = lat.numpy()
    out["latent_packed"] = packer(torch.sign(lat)).numpy()
    np.savez_compressed(HERE / "pack_kat.npz", **out)
    print("pack_kat.npz", {k: v.shape for k, v in out.items()})


SMALL_CASES = [  # (seed, K, N, M, bias)
    (1, 64, 24, 1, False), (2, 128, 40, 3, False), (3, 256, 256, 8, True), (4, 512, 96, 5, False),
    (5, 1024, 72, 2, False), (6, 8, 4, 1, False), (7, 136, 33, 7, True), (8, 4096, 64, 4, False),
]
FULL_M = (1, 3)


def gen_forward(bitnet):
    out = {}
    meta = []
    for seed, k, n, m, bias in SMALL_CASES:
        case = oracle.synth_case(seed, k, n, m, with_bias=bias)
        out[f"small_{seed}_y"] = ref_forward(bitnet, case)
        meta.append((seed, k, n, m, int(bias)))
    out["small_meta"] = np.array(meta, dtype=np.int64)
    # 3-D input [B, T, K] (reference broadcasting through .view(1, K))
    case = oracle.synth_case(21, 256, 48, 6)
    y3 = ref_forward(bitnet, {**case, "x": case["x"].reshape(2, 3, 256)})
    out["x3d_21_y"] = y3
    # full LLaMA shapes, fp32 reference
    fmeta = []
    for idx, (name, (k, n)) in enumerate(oracle.LLAMA_SHAPES.items()):
        for m in FULL_M:
            seed = 100 + 10 * idx + m
            case = oracle.synth_case(seed, k, n, m)
            out[f"full_{name}_m{m}_y"] = ref_forward(bitnet, case)
            fmeta.append((seed, k, n, m))
            print("full", name, m, flush=True)
    out["full_meta"] = np.array(fmeta, dtype=np.int64)
    out["full_names"] = np.array([f"{n}_m{m}" for n in oracle.LLAMA_SHAPES for m in FULL_M])
    # the reference's own fp16 path on one shape (documents its distance from fp32, H4)
    case = oracle.synth_case(300, 4096, 4096, 3)
    out["fp16_7b_attn_m3_y_fp16path"] = ref_forward(bitnet, case, torch.float16)
    out["fp16_7b_attn_m3_y_fp32path"] = ref_forward(bitnet, case, torch.float32)
    np.savez_compressed(HERE / "forward_golden.npz", **out)
    print("forward_golden.npz written")


def gen_train_inf_equiv(bitnet, packer):
    """BitLinear (training, fp latent weight) vs BitLinearInf (packed) — SURVEY §4 implicit equivalence."""
    torch.manual_seed(7)
    k, n, m = 256, 64, 5
    lat = torch.randn(n, k)
    train = bitnet.BitLinear(k, n, dtype=torch.float32)
    inf = bitnet.BitLinearInf(k, n, dtype=torch.float32)
    g = torch.rand(n) + 0.5
    h = torch.rand(k) * 3 - 1.5
    with torch.no_grad():
        train.weight.copy_(lat)
        train.weight_scale.copy_(g)
        train.input_factor.copy_(h)
        inf.weight.copy_(packer(torch.sign(lat)))
        inf.weight_scale.copy_(g)
        inf.input_factor.copy_(h)
        x = torch.randn(m, k)
        yt = train(x)
        yi = inf(x)
    np.savez_compressed(HERE / "train_inf_equiv.npz", latent=lat.numpy(), g=g.numpy(), h=h.numpy(), x=x.numpy(),
                        y_train=yt.numpy(), y_inf=yi.numpy(), packed=inf.weight.numpy())
    print("train_inf_equiv max diff", float((yt - yi).abs().max()))


def import_reference_transformers():
    stub = types.ModuleType("transformers.dependency_versions_check")
    stub.dep_version_check = lambda *a, **k: None
    sys.modules["transformers.dependency_versions_check"] = stub
    sys.path.insert(0, str(REF / "transformers/src"))
    import transformers  # noqa: F401  (the reference's vendored 4.35.0.dev0)
    from transformers import BitLlamaConfig, BitLlamaForCausalLMInf
    return BitLlamaConfig, BitLlamaForCausalLMInf


TINY = dict(vocab_size=384, hidden_size=256, intermediate_size=768, num_hidden_layers=2, num_attention_heads=2,
            num_key_value_heads=2, max_position_embeddings=256, rms_norm_eps=1e-6, hidden_act="silu",
            rope_theta=10000.0, pad_token_id=0, bos_token_id=1, eos_token_id=None, tie_word_embeddings=False)


def gen_model():
    """Tiny seeded BitLlamaForCausalLMInf: full-sequence logits, the lm_eval.py:93-124 perplexity on a fixed
    slice, and greedy decode tokens (generation/utils.py greedy_search), all fp32 on CPU."""
    BitLlamaConfig, BitLlamaForCausalLMInf = import_reference_transformers()
    cfg = BitLlamaConfig(**TINY)
    torch.manual_seed(0)
    model = BitLlamaForCausalLMInf(cfg).float().eval()
    rng = np.random.Generator(np.random.PCG64(2024))
    sd = model.state_dict()
    with torch.no_grad():
        for name, p in sd.items():
            if name.endswith("_proj.weight"):
                p.copy_(torch.from_numpy(rng.integers(-128, 128, size=tuple(p.shape), dtype=np.int8)))
            elif name.endswith("weight_scale"):
                p.copy_(torch.from_numpy(rng.uniform(0.5, 1.5, size=tuple(p.shape)).astype(np.float16)).float())
            elif name.endswith("input_factor"):
                p.copy_(torch.from_numpy(rng.uniform(-1.5, 1.5, size=tuple(p.shape)).astype(np.float16)).float())
            elif "layernorm.weight" in name or name == "model.norm.weight":
                p.copy_(torch.from_numpy(rng.uniform(0.5, 1.5, size=tuple(p.shape)).astype(np.float16)).float())
            elif name in ("model.embed_tokens.weight", "lm_head.weight"):
                std = 0.5 if name.startswith("model.") else 0.05   # keeps logits O(1) so the PPL is well conditioned
                p.copy_(torch.from_numpy((rng.standard_normal(size=tuple(p.shape)) * std).astype(np.float16)).float())
    ids = torch.from_numpy(rng.integers(3, cfg.vocab_size, size=(2, 48), dtype=np.int64))
    with torch.no_grad():
        out = model(ids)
        logits = out.logits.float()
        # lm_eval.py:99-124 — hidden -> lm_head -> shift -> CE(mean) * seqlen, exp(sum / (nsamples*seqlen))
        nlls = []
        seqlen = ids.shape[1]
        for i in range(ids.shape[0]):
            hs = model.model(ids[i:i + 1])[0]
            lg = model.lm_head(hs)
            shift_logits = lg[:, :-1, :]
            shift_labels = ids[i:i + 1][:, 1:]
            loss = torch.nn.CrossEntropyLoss()(shift_logits.reshape(-1, shift_logits.size(-1)), shift_labels.reshape(-1))
            nlls.append(loss.float() * seqlen)
        ppl = torch.exp(torch.stack(nlls).sum() / (ids.shape[0] * seqlen))
        prompt = ids[:, :8]
        gen = model.generate(prompt, max_new_tokens=24, do_sample=False, eos_token_id=None, pad_token_id=0)
    fix = {f"sd::{k}": v.numpy() for k, v in model.state_dict().items()}
    fix.update(input_ids=ids.numpy(), logits=logits.numpy(), ppl=np.float64(ppl.item()), prompt=prompt.numpy(),
               generated=gen.numpy(), config_keys=np.array(list(TINY.keys())),
               config_vals=np.array([str(v) for v in TINY.values()]))
    np.savez_compressed(HERE / "tiny_model.npz", **fix)
    print("tiny_model.npz ppl", ppl.item(), "generated", gen.shape)


def gen_ppl512():
    """A fixed 512-token slice for the perplexity-parity gate (north_star: "perplexity on a fixed 512-token slice
    matches to 3 decimals"). The slice is SAMPLED from the tiny reference model itself (temperature 0.1), so
    the perplexity is in the range of a trained model instead of ~vocab_size; the number stored is the reference's:
    evaluation/lm_eval.py:99-124 on one 512-token window, fp32 on CPU. Weights are those of tiny_model.npz."""
    BitLlamaConfig, BitLlamaForCausalLMInf = import_reference_transformers()
    z = np.load(HERE / "tiny_model.npz")
    cfg = BitLlamaConfig(**dict(TINY, max_position_embeddings=512))
    model = BitLlamaForCausalLMInf(cfg).float().eval()
    model.load_state_dict({k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd::")})
    torch.manual_seed(20240917)
    seqlen = 512
    prompt = torch.tensor([[1, 17, 230, 5]], dtype=torch.int64)
    with torch.no_grad():
        ids = model.generate(prompt, max_new_tokens=seqlen - prompt.shape[1], do_sample=True, temperature=0.1, top_k=0,
                             eos_token_id=None, pad_token_id=0)
        assert ids.shape == (1, seqlen)
        hs = model.model(ids)[0]
        lg = model.lm_head(hs)
        loss = torch.nn.CrossEntropyLoss()(lg[:, :-1, :].reshape(-1, lg.size(-1)), ids[:, 1:].reshape(-1))
        nll = loss.float() * seqlen
        ppl = torch.exp(nll / seqlen)
        # the same quantity accumulated in float64 from the fp32 logits (what the parity test recomputes)
        loss64 = torch.nn.functional.cross_entropy(lg[0, :-1].double(), ids[0, 1:])
        ppl64 = torch.exp(loss64 * seqlen / seqlen)
    np.savez_compressed(HERE / "tiny_ppl512.npz", input_ids=ids.numpy(), ppl=np.float64(ppl.item()), ppl64=np.float64(ppl64.item()),
                        last_logits=lg[0, -1].float().numpy(), logits_stride64=lg[0, ::64].float().numpy())
    print("tiny_ppl512.npz ppl", ppl.item(), "ppl64", ppl64.item())


def gen_rope():
    """cos/sin caches of the reference's LlamaRotaryEmbedding (modeling_bitllama.py:87-121), head_dim 128."""
    import_reference_transformers()
    from transformers.models.bitllama.modeling_bitllama import LlamaRotaryEmbedding
    rot = LlamaRotaryEmbedding(128, max_position_embeddings=512, base=10000.0)
    cos, sin = rot(torch.zeros(1, 1, 512, 128), seq_len=512)
    np.savez_compressed(HERE / "rope_tables.npz", cos=cos.float().numpy(), sin=sin.float().numpy())
    print("rope_tables.npz", cos.shape)


if __name__ == "__main__":
    torch.set_grad_enabled(False)
    which = set(sys.argv[1:]) or {"pack", "forward", "equiv", "model", "rope", "ppl512"}
    bitnet = load_bitnet()
    packer = load_packer()
    if "pack" in which:
        gen_pack(bitnet, packer)
    if "equiv" in which:
        gen_train_inf_equiv(bitnet, packer)
    if "forward" in which:
        gen_forward(bitnet)
    if "model" in which:
        gen_model()
    if "rope" in which:
        gen_rope()
    if "ppl512" in which:
        gen_ppl512()
