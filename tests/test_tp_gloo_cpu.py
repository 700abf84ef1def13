"""world_size = 2 gloo test (CPU) of the tensor-parallel host logic: `onebit_b200.tp.shard_state_dict` + the two
collectives a sharded BitLinear needs, with the CPU oracle standing in for the CUDA kernels on each rank.
Column-parallel (q_proj): all-reduce of per-token (sum, sumsq). Row-parallel (o_proj): all-reduce of partial sums."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from onebit_b200.tp import pad_to, shard_state_dict
from oracle import oracle


def _tiny(golden_dir):
    z = np.load(golden_dir / "tiny_model.npz")
    cfg = {k: v for k, v in zip(z["config_keys"], z["config_vals"])}
    config = {k: int(cfg[k]) for k in ("hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads",
                                       "vocab_size")}
    sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd::")}
    return config, sd


def _worker(rank, world, port, golden_dir, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        config, sd = _tiny(golden_dir)
        H = config["hidden_size"]
        shard = shard_state_dict(config, sd, world, rank)
        rng = np.random.Generator(np.random.PCG64(11))
        x = rng.standard_normal((3, H)).astype(np.float32)
        pre = "model.layers.0.self_attn."
        # ---- column-parallel q_proj: local rows, global LayerNorm statistics through an all-reduce
        wq, gq, hq = (shard[pre + "q_proj." + k].numpy() for k in ("weight", "weight_scale", "input_factor"))
        _, u = oracle.bitlinear_forward_c(x, wq, gq, hq, return_pre_ln=True)
        stats = torch.from_numpy(np.stack([u.sum(-1), (u.astype(np.float64) ** 2).sum(-1)], -1).astype(np.float64))
        dist.all_reduce(stats)
        mean = (stats[:, 0] / H).numpy()[:, None]
        var = (stats[:, 1] / H).numpy()[:, None] - mean ** 2
        y_local = (u - mean) / np.sqrt(var + 1e-5)
        full = oracle.bitlinear_forward_c(x, sd[pre + "q_proj.weight"].numpy(), sd[pre + "q_proj.weight_scale"].numpy(),
                                          sd[pre + "q_proj.input_factor"].numpy())
        Hl = H // world
        err_col = oracle.rel_l2(y_local, full[:, rank * Hl:(rank + 1) * Hl])
        # ---- row-parallel o_proj: local (zero-padded) K slice, partial sums all-reduced, then scale + LayerNorm
        wo, go, ho = (shard[pre + "o_proj." + k].numpy() for k in ("weight", "weight_scale", "input_factor"))
        Hk = pad_to(Hl)
        assert wo.shape == (H, Hk // 8) and ho.shape == (Hk,) and (ho[Hl:] == 0).all()
        xl = np.zeros((3, Hk), dtype=np.float32)
        xl[:, :Hl] = x[:, rank * Hl:(rank + 1) * Hl]
        _, part = oracle.bitlinear_forward_c(xl, wo, np.ones(H, np.float32), ho, return_pre_ln=True)
        t = torch.from_numpy(part.astype(np.float32))
        dist.all_reduce(t)
        y = oracle.layernorm_np(t.numpy() * go[None, :])
        full_o = oracle.bitlinear_forward_c(x, sd[pre + "o_proj.weight"].numpy(), sd[pre + "o_proj.weight_scale"].numpy(),
                                            sd[pre + "o_proj.input_factor"].numpy())
        err_row = oracle.rel_l2(y, full_o)
        q.put((rank, err_col, err_row))
    finally:
        dist.destroy_process_group()


def test_tp_shards_and_collectives_world2(golden_dir):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + (os.getpid() % 150)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, golden_dir, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err_col, err_row in out:
        assert err_col < 1e-5, (rank, err_col)
        assert err_row < 1e-5, (rank, err_row)


def test_shard_shapes_for_llama_sizes():
    # shapes only (meta tensors): every LLaMA-7B/13B tp in {2,4,8} produces 256-column-aligned row-parallel shards
    for H, I, heads in [(4096, 11008, 32), (5120, 13824, 40)]:
        for tp in (2, 4, 8):
            config = dict(hidden_size=H, intermediate_size=I, num_attention_heads=heads)
            sd = {"model.layers.0.self_attn.o_proj.weight": torch.zeros((8, H // 8), dtype=torch.int8),
                  "model.layers.0.self_attn.o_proj.input_factor": torch.ones(H),
                  "model.layers.0.mlp.down_proj.weight": torch.zeros((8, I // 8), dtype=torch.int8),
                  "model.layers.0.mlp.down_proj.input_factor": torch.ones(I),
                  "model.layers.0.mlp.gate_proj.weight": torch.zeros((I, 8), dtype=torch.int8),
                  "model.layers.0.mlp.gate_proj.weight_scale": torch.ones(I)}
            out = shard_state_dict(config, sd, tp, tp - 1)
            assert out["model.layers.0.self_attn.o_proj.weight"].shape[1] * 8 == pad_to(H // tp)
            assert out["model.layers.0.mlp.down_proj.weight"].shape[1] * 8 == pad_to(I // tp)
            assert out["model.layers.0.mlp.down_proj.input_factor"].shape[0] == pad_to(I // tp)
            assert out["model.layers.0.mlp.gate_proj.weight"].shape[0] == I // tp
            assert out["model.layers.0.mlp.gate_proj.weight_scale"].shape[0] == I // tp
