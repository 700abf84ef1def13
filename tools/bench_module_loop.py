"""The reference-style module loop on the GPU: 224 `BitLinearB200.forward` calls per token (7 projections x 32 layers of
LLaMA-7B, bitnet.py:112-122 as the drop-in module executes it: quantise | GEMV | scale + LayerNorm, three launches per
call) timed eagerly and as one CUDA-graph replay, beside the fused decoder's four stage launches per layer.
Only the BitLinear calls are issued (attention, norms and residuals of the model are left out), each on its own weights
(> L2 per token). CUDA events on the launching stream.

    python tools/bench_module_loop.py [--batch 1]
"""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from onebit_b200 import BitLinearB200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--layers", type=int, default=32)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    H, I, L, M = 4096, 11008, args.layers, args.batch
    gen = torch.Generator(device=dev).manual_seed(0)

    def mod(k, n):
        m = BitLinearB200(k, n, device=dev, dtype=torch.float16)
        with torch.no_grad():
            m.weight.copy_(torch.randint(-128, 128, (n, k // 8), dtype=torch.int8, device=dev, generator=gen))
            m.weight_scale.copy_(torch.rand(n, device=dev, generator=gen) + 0.5)
            m.input_factor.copy_(torch.rand(k, device=dev, generator=gen) * 3 - 1.5)
        return m

    layers = [dict(q=mod(H, H), k=mod(H, H), v=mod(H, H), o=mod(H, H), gate=mod(H, I), up=mod(H, I), down=mod(I, H))
              for _ in range(L)]
    x = torch.randn(M, H, device=dev, generator=gen).half()

    def token(x):
        for ly in layers:  # the data dependencies of a decoder layer, BitLinear calls only
            q, k, v = ly["q"](x), ly["k"](x), ly["v"](x)
            o = ly["o"](q + k + v)
            g, u = ly["gate"](o), ly["up"](o)
            x = ly["down"](g * u)
        return x

    with torch.no_grad():
        for _ in range(3):
            token(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            token(x)
        e1.record()
        torch.cuda.synchronize()
        eager_ms = e0.elapsed_time(e1) / reps
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            token(x)
            s.synchronize()
            with torch.cuda.graph(g, stream=s):
                token(x)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        graph_ms = e0.elapsed_time(e1) / reps
    calls = 7 * L
    print(json.dumps({"what": "reference-style module loop, BitLinear calls only", "model": "LLaMA-7B shapes", "layers": L,
                      "batch": M, "calls_per_token": calls,
                      "eager_ms_per_token": round(eager_ms, 3), "eager_us_per_call": round(eager_ms * 1e3 / calls, 2),
                      "graph_ms_per_token": round(graph_ms, 3), "graph_us_per_call": round(graph_ms * 1e3 / calls, 2),
                      "note": "graph replay also holds the 3 elementwise torch kernels per layer that stand in for the glue"}))


if __name__ == "__main__":
    main()
