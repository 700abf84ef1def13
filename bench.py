#!/usr/bin/env python
"""bench.py — LLaMA-7B-OneBit greedy decode throughput on B200 (BASELINE.json configs[1]) + roofline of the
dominant kernel (bit-plane IMMA packed GEMV) + the reference's CPU path timed beside it.

    python bench.py --gpus 1 --steps 128 --warmup 8
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1      # reference CPU arm (rank 0 only)

A "step" is one decode step (one new token per sequence) of the full model: embedding, 32 x (RMSNorm, q/k/v/o and
gate/up/down BitLinear with their LayerNorms, RoPE, attention over the KV cache, SiLU*up, residuals), final norm,
fp16 lm_head, greedy argmax. Weights are synthetic (random packed signs, random g/h, N(0, 0.02) embeddings); every
step streams 810 MB of packed signs + 262 MB of lm_head, far more than the 126 MB L2, so no explicit L2 flush.
Multi-GPU = independent replicas (the path does not shard below one model replica at these batch sizes; weak
scaling), no data-path collective; time = max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "LLaMA-7B-OneBit decode tok/s (greedy, batch 1 per GPU)"  # BASELINE.json metric, configs[1]
UNIT = "tok/s"


def model_config(name: str):
    from onebit_b200 import LLAMA2_13B, LLAMA_7B
    return {"7b": ("LLaMA-7B-OneBit", LLAMA_7B), "13b": ("LLaMA2-13B-OneBit", LLAMA2_13B)}[name]


def workload_config(name: str, B: int, world: int, tp: bool, tp_allreduce=None, prompt_len: int = 16) -> dict:
    """The `config` object of the JSON line — the same for both arms (the reference arm runs this workload's op sequence
    on the host cores; what it samples is said in its `cpu_baseline.sample` / `note`)."""
    return {"workload": f"{name} greedy decode, batch {B} per GPU, {prompt_len}-token prompt then greedy decode steps, "
                        f"static KV cache, one CUDA-graph replay per step",
            "replicas": 1 if (tp and world > 1) else world,
            "parallelism": (f"tp{world}: one tensor-parallel replica, 4 all-reduces per layer ({tp_allreduce}: "
                            "one-shot Lamport all-reduce over NVLink peer memory, csrc/p2p_allreduce.cu)"
                            if (tp and world > 1) else f"{world} independent replica(s), no data-path collective"),
            "l2": "per-step weight stream 1.07 GB > 126 MB L2 (no flush needed)",
            "activation_dtype": ("fp32 residual / fp16 KV cache / 23-bit integer BitLinear inputs" if B <= 4 else
                                 "fp32 residual / fp16 KV cache / fp16 BitLinear inputs (tcgen05 kind::f16, fp32 accumulate)")}


def bitlinear_bytes(cfg, batch: int) -> dict:
    """Algorithmic bytes: packed signs + fp16 x/y + fp16 g/h per BitLinear call."""
    H, I, L = cfg["hidden_size"], cfg["intermediate_size"], cfg["num_hidden_layers"]
    shapes = [(H, H)] * 4 + [(H, I)] * 2 + [(I, H)]
    per_layer = sum(n * k // 8 + 2 * batch * k + 2 * batch * n + 2 * (n + k) for k, n in shapes)
    packed = sum(n * k // 8 for k, n in shapes)
    return {"per_step": per_layer * L, "packed_per_step": packed * L, "gemv_launches": 4 * L}


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.path = Path(f"/tmp/onebit_clocks_{os.getpid()}.csv")

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                          "50", "-i", str(self.index)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = []
        for line in self.path.read_text().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                try:
                    rows.append((float(parts[0]), float(parts[1]), float(parts[2]), parts[3:7]))
                except ValueError:
                    pass
        self.path.unlink(missing_ok=True)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = sorted(r[0] for r in rows if r[2] > 200.0) or sorted(r[0] for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows for i, v in enumerate(r[3]) if v.lower().startswith("active")})
        return {"sm_mhz": busy[len(busy) // 2], "sm_max_mhz": rows[0][1], "power_w_max": max(r[2] for r in rows),
                "samples": len(rows), "reasons": reasons}


# ------------------------------------------------------------------------------------------------
# reference CPU arm / cpu_baseline (oracle/: allowed here as the baseline being timed, never as product)
# ------------------------------------------------------------------------------------------------
def cpu_decode_baseline(cfg, batch: int, tokens: int, threads: int | None = None) -> dict:
    """Times the op-for-op port of the reference's CPU path (oracle/ref_port.py) on ONE decoder layer for `tokens`
    decode steps + one lm_head call, and extrapolates to the full model. Returns tok/s."""
    import torch
    from onebit_b200 import synthetic_state_dict
    from oracle import ref_port
    if threads:
        torch.set_num_threads(threads)
    one = dict(cfg, num_hidden_layers=1)
    sd = synthetic_state_dict(one, seed=0, param_dtype=torch.float32)
    model = ref_port.RefPortModel(one, sd, dtype=torch.float32)
    H = cfg["hidden_size"]
    gen = torch.Generator().manual_seed(0)
    past = (torch.randn(batch, cfg["num_attention_heads"], 16, H // cfg["num_attention_heads"], generator=gen),) * 2
    hs = torch.randn(batch, 1, H, generator=gen)
    lm = sd["lm_head.weight"].float()
    with torch.no_grad():
        model.layer(0, hs, 16, past)  # warm-up
        t0 = time.perf_counter()
        for i in range(tokens):
            model.layer(0, hs, 16 + i, past)
        t_layer = (time.perf_counter() - t0) / tokens
        t0 = time.perf_counter()
        torch.nn.functional.linear(hs, lm)
        t_head = time.perf_counter() - t0
    step_s = t_layer * cfg["num_hidden_layers"] + t_head
    return {"value": batch / step_s, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"oracle/ref_port.py (op-for-op torch-CPU port of bitnet.py:112-122 + decoder layer), fp32, "
                      f"1 decoder layer x {tokens} decode steps ({t_layer * 1e3:.0f} ms/layer) + 1 lm_head call, "
                      f"extrapolated x{cfg['num_hidden_layers']} layers; host cpu_count={os.cpu_count()}",
            "ms_per_step_extrapolated": step_s * 1e3}


def run_reference(args, rank: int):
    if rank != 0:
        return
    name, cfg = model_config(args.model)
    res = None
    t_all = time.perf_counter()
    vals = []  # each "step" is one bounded sample of the layer-level workload (the sampler warms itself up)
    for _ in range(max(1, args.steps)):
        res = cpu_decode_baseline(cfg, args.batch, tokens=2)
        vals.append(res["value"])
        if time.perf_counter() - t_all > 150:
            break
    v = sum(vals) / len(vals)
    res["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals),
            "warmup": args.warmup, "ms_per_step": 1e3 * args.batch / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(name, args.batch, max(1, args.gpus), False),
            "note": "reference CPU path: the reference's op sequence for this workload (torch ATen ops, fp32) timed on the host "
                    "cores over one decoder layer + lm_head and EXTRAPOLATED to the model's layer count (cpu_baseline.sample)",
            "cpu_baseline": res, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist
    from onebit_b200 import BitLlamaDecoderB200, _lib, synthetic_state_dict

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = _lib.load()
    _lib.check(lib.onebit_device_check(local_rank), "onebit_device_check")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    name, cfg = model_config(args.model)
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        pk = json.loads(peaks_path.read_text())
        peak, peak_src = pk["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        tpeak, tpeak_src = pk.get("bf16_tflops_sustained", pk.get("bf16_tflops", 1590.0)), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        tpeak, tpeak_src = 1400.0, "fallback (B200_PROFILING.md sustained)"

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    from onebit_b200 import replicas

    def max_over_ranks(ms: float) -> float:
        return replicas.max_over_ranks(ms, device=dev)

    def measure(model_name, cfg, B, K, W, sampler=None, with_e2e=True):
        """Decode throughput of one replica per GPU at batch B: device-resident loop, end-to-end loop with host ids in and
        out every step, and the BitLinear projection launches of a step on their own (roofline of the dominant kernel)."""
        prompt_len = 16
        tp = bool(args.tp) and world > 1
        sd = synthetic_state_dict(cfg, seed=0 if tp else rank)  # (a tensor-parallel replica: every rank shards the same model)
        dec = BitLlamaDecoderB200(cfg, sd, device=dev, max_seq_len=prompt_len + 2 * (K + W) + 32, max_batch=B,
                                  tp_group=dist.group.WORLD if tp else None)
        del sd
        nrep = 1 if tp else world  # replicas whose tokens add up
        gen = torch.Generator().manual_seed(1234 + (0 if tp else rank))
        prompt = torch.randint(3, cfg["vocab_size"], (B, prompt_len), generator=gen)

        def prefill():
            dec.reset(prompt[:, 0])
            ids = prompt.to(dev)
            for i in range(prompt_len):
                dec.step(ids[:, i])

        # ---- device-resident decode: K timed steps
        prefill()
        for _ in range(max(W, 3)):
            dec.step()
        if sampler:
            sampler.start()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            dec.step()
        e1.record()
        barrier()
        if tp:  # one replica spread over the ranks: its tokens count once, over the slowest rank's time
            ms = max_over_ranks(e0.elapsed_time(e1))
            value = B * K / (ms * 1e-3)
        else:   # independent replicas: all tokens of all ranks over the slowest rank's time
            agg = replicas.aggregate_throughput(B * K, e0.elapsed_time(e1), device=dev)
            ms, value = agg["ms"], agg["per_s"]
        out = {"batch": B, "value": value, "ms_per_step": ms / K, "launches_per_step": dec.launches_per_step(),
               "prompt_len": prompt_len, "tp_allreduce": getattr(dec, "tp_allreduce", "none")}
        # ---- end to end through the public API: ids from pinned host memory in, next ids back to the host, every step
        if with_e2e:
            prefill()
            h_in = torch.empty(B, dtype=torch.int64).pin_memory()
            h_out = torch.empty(B, dtype=torch.int64).pin_memory()
            h_in.copy_(dec.next_ids().cpu())
            for _ in range(max(W, 3)):
                dec.step(h_in)
                h_out.copy_(dec.next_ids(), non_blocking=True)
                torch.cuda.current_stream(dev).synchronize()
                h_in.copy_(h_out)
            barrier()
            e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e2.record()
            for _ in range(K):
                dec.step(h_in)                                   # H2D of the fed ids (pinned -> device) + graph replay
                h_out.copy_(dec.next_ids(), non_blocking=True)   # D2H of the step's result
                torch.cuda.current_stream(dev).synchronize()
                h_in.copy_(h_out)                                # host feeds the token back (what a serving loop does)
            e3.record()
            barrier()
            ms_e2e = max_over_ranks(e2.elapsed_time(e3))
            out["e2e"] = {"value": nrep * B * K / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 8 * B,
                          "d2h_bytes_per_step": 8 * B, "ms_per_step": ms_e2e / K}
        if tp:  # (the per-kernel roofline is reported by the single-GPU runs; a tensor-parallel step is collective-latency-bound)
            dec.close()
            del dec
            torch.cuda.empty_cache()
            return out
        # ---- dominant kernel alone: the 4 x L BitLinear projection launches of a step, back to back, CUDA events
        bb = bitlinear_bytes(cfg, B)
        g = torch.cuda.CUDAGraph()
        _lib.check(lib.onebit_decoder_gemv_only(dec._handle, B, torch.cuda.current_stream(dev).cuda_stream), "gemv_only")
        torch.cuda.synchronize(dev)
        with torch.cuda.graph(g):
            _lib.check(lib.onebit_decoder_gemv_only(dec._handle, B, torch.cuda.current_stream(dev).cuda_stream), "gemv_only")
        for _ in range(3):
            g.replay()
        barrier()
        reps = 20 if B <= 4 else 5
        e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e4.record()
        for _ in range(reps):
            g.replay()
        e5.record()
        barrier()
        gemv_ms = e4.elapsed_time(e5) / reps
        launches = bb["gemv_launches"]
        bytes_per_launch = bb["per_step"] / launches
        achieved = bytes_per_launch / (gemv_ms * 1e-3 / launches) / 1e9
        H, I, L = cfg["hidden_size"], cfg["intermediate_size"], cfg["num_hidden_layers"]
        flops_step = 2.0 * B * L * (4 * H * H + 3 * H * I)
        if B > 4:
            kname = ("prefill_tc5_kernel<decode tile> (tcgen05.mma kind::f16, packed signs expanded in registers with "
                     "input_factor folded in, TMA activations, split-K; one launch per projection group)")
        elif os.environ.get("ONEBIT_FUSED", "1") != "0" and os.environ.get("ONEBIT_FUSED_V2", "1") != "0" and B <= 2:
            kname = ("fused2::fused_gemv2_kernel (glue prologue from the producer's records, no block reduction + bit-plane "
                     "IMMA packed-sign GEMV, one launch per BitLinear group)")
        elif os.environ.get("ONEBIT_FUSED", "1") != "0":
            kname = "fused::fused_gemv_kernel (glue prologue + bit-plane IMMA packed-sign GEMV, one launch per BitLinear group)"
        else:
            kname = "imma::gemv_kernel (bit-plane IMMA packed-sign GEMV)"
        traffic = None
        tpath = ROOT / "profiles" / ("r02_fused2_traffic.json" if "fused2" in kname else "r01_gemv_traffic.json")
        if tpath.exists() and B <= 4:
            traffic = json.loads(tpath.read_text()).get("dram_bytes_per_launch")
        out["roofline"] = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                           "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                           "bytes_per_launch": bytes_per_launch, "us_per_launch": gemv_ms * 1e3 / launches,
                           "launches_timed": launches * reps,
                           "how": f"CUDA events around {reps} replays of a graph holding the step's 4xL BitLinear projection "
                                  "launches (attention / glue / lm_head left out), distinct weights per launch "
                                  "(weight stream per replay > L2)",
                           "tensor": {"achieved_tflops": flops_step / (gemv_ms * 1e-3) / 1e12, "peak_tflops": tpeak,
                                      "frac": flops_step / (gemv_ms * 1e-3) / 1e12 / tpeak, "peak_source": tpeak_src},
                           "step_level": {"bitlinear_GBs_over_whole_step": bb["per_step"] / (ms / K * 1e-3) / 1e9,
                                          "frac": bb["per_step"] / (ms / K * 1e-3) / 1e9 / peak}}
        dec.close()
        del dec
        torch.cuda.empty_cache()
        return out

    B, K, W = args.batch, args.steps, args.warmup
    metric = METRIC if (args.model == "7b" and B == 1) else f"{name} decode tok/s (greedy, batch {B} per GPU)"
    sampler = ClockSampler(local_rank) if rank == 0 else None
    main_res = measure(name, cfg, B, K, W, sampler=sampler)
    clocks = sampler.stop() if sampler else None
    # the metric's second operating point ("decode tok/s @ b=1/32"): the batched tcgen05 path, same model, same run
    extra = {}
    if B == 1 and os.environ.get("ONEBIT_BENCH_EXTRA", "1") != "0":
        for label, mname, eb in (("batch32", args.model, 32),):
            try:
                ename, ecfg = model_config(mname)
                r = measure(ename, ecfg, eb, min(K, 32), 3, with_e2e=True)
                r["workload"] = f"{ename} greedy decode, batch {eb} per GPU, 16-token prompt then {min(K, 32)} generated tokens"
                extra[label] = r
            except Exception as exc:  # the headline line must survive a failure of the secondary operating point
                extra[label] = {"error": str(exc)[:300]}
    if B == 1 and world == 1 and os.environ.get("ONEBIT_BENCH_EXTRA", "1") != "0":
        # BASELINE configs[2]: prompt pass 2048 tokens x batch 8 through the tcgen05 prefill tile, then 128 decode steps
        try:
            r = subprocess.run([sys.executable, str(ROOT / "tools" / "bench_config3.py"), "--model", args.model], capture_output=True,
                               text=True, timeout=400)
            extra["prefill2048_decode128_b8"] = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else \
                {"error": r.stderr[-300:]}
        except Exception as exc:
            extra["prefill2048_decode128_b8"] = {"error": str(exc)[:300]}
    if world > 1:
        # (no destroy_process_group: it blocks while CUDA graphs that captured collective kernels exist; see tools/tp_check.py)
        torch.cuda.synchronize(dev)
        dist.barrier()
    if rank != 0:
        sys.stdout.flush()
        os._exit(0)
    cpu = cpu_decode_baseline(cfg, B, tokens=3) if world == 1 else None
    line = {"metric": metric, "value": main_res["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": main_res["ms_per_step"], "higher_is_better": True,
            "scaling": "strong" if (args.tp and world > 1) else "weak", "vs_baseline": None,
            "dtype": "int8" if B <= 4 else "f16", "data": "synthetic",
            "config": workload_config(name, B, world, bool(args.tp), main_res.get("tp_allreduce"), main_res["prompt_len"]),
            "e2e": main_res["e2e"],
            "gpu_launches": main_res["launches_per_step"] * K, "launches_per_step": main_res["launches_per_step"],
            "roofline": main_res.get("roofline"), "clocks": clocks}
    line.update(extra)
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if world > 1:
        sys.stdout.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("ONEBIT_BENCH_BATCH", "1")))
    ap.add_argument("--model", default=os.environ.get("ONEBIT_BENCH_MODEL", "7b"), choices=["7b", "13b"])
    ap.add_argument("--tp", action="store_true", default=os.environ.get("ONEBIT_BENCH_TP", "0") == "1",
                    help="ONE tensor-parallel replica over all ranks (BASELINE configs[4]) instead of independent replicas")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch multi-GPU runs with torch.distributed.run (one rank per GPU)")
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
