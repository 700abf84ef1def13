"""Kernel-level timing of the BitLinear forward over a rotating set of distinct weight matrices (> L2),
per LLaMA projection shape and batch size. CUDA-event timing on the launching stream.

    python tools/bench_layers.py [--variants simt,mma] [--ms 1,8,32] [--graph]
"""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import onebit_b200  # noqa: E402
from onebit_b200 import _lib  # noqa: E402

SHAPES = {"7b_attn": (4096, 4096), "7b_gate_up": (4096, 11008), "7b_down": (11008, 4096),
          "13b_attn": (5120, 5120), "13b_gate_up": (5120, 13824), "13b_down": (13824, 5120)}


def alg_bytes(k, n, m, a=2, p=2):
    return n * k // 8 + a * m * k + a * m * n + p * (n + k)


def time_shape(k, n, m, variant, use_graph, peak):
    dev = torch.device("cuda:0")
    nmat = max(4, int(400e6 // (n * k // 8)))  # > 3x L2 of distinct weights
    ws = [torch.randint(-128, 128, (n, k // 8), dtype=torch.int8, device=dev) for _ in range(nmat)]
    g = (torch.rand(n, device=dev) + 0.5).half()
    h = (torch.rand(k, device=dev) * 3 - 1.5).half()
    x = torch.randn(m, k, device=dev).half()
    lib = _lib.load()
    t = torch.empty(m, n, dtype=torch.float32, device=dev)
    wsb = lib.onebit_bitlinear_workspace_bytes(m, k, n)
    wsp = torch.empty(wsb, dtype=torch.uint8, device=dev)
    mwsb = lib.onebit_matvec_workspace_bytes(m, k)
    y = torch.empty(m, n, dtype=torch.float16, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    var = _lib.VARIANTS[variant]

    def launch_all(which):
        for w in ws:
            if which == "matvec":
                rc = lib.onebit_bitlinear_matvec(x.data_ptr(), w.data_ptr(), g.data_ptr(), h.data_ptr(), t.data_ptr(), m,
                                                 k, n, 0, 0, 0, wsp.data_ptr(), mwsb, var, stream)
            else:
                rc = lib.onebit_bitlinear_forward(x.data_ptr(), w.data_ptr(), g.data_ptr(), h.data_ptr(), None,
                                                  y.data_ptr(), m, k, n, 0, 0, 1e-5, wsp.data_ptr(), wsb, var, stream)
            assert rc == 0, _lib.last_error()

    out = {}
    for which in ("matvec", "forward"):
        if use_graph:
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                stream = s.cuda_stream
                launch_all(which)
                s.synchronize()
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, stream=s):
                    stream = torch.cuda.current_stream().cuda_stream
                    launch_all(which)
            run = gr.replay
            stream = torch.cuda.current_stream().cuda_stream
        else:
            run = lambda: launch_all(which)  # noqa: E731
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (reps * nmat)
        gbs = alg_bytes(k, n, m) / us / 1e3
        out[which] = {"us": round(us, 3), "GBs": round(gbs, 1), "frac": round(gbs / peak, 4)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", default="simt")
    ap.add_argument("--ms", default="1")
    ap.add_argument("--shapes", default=",".join(SHAPES))
    ap.add_argument("--graph", action="store_true")
    args = ap.parse_args()
    peaks = json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text()) \
        if (Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6650.0}
    peak = peaks["hbm_gbs"]
    for variant in args.variants.split(","):
        for name in args.shapes.split(","):
            k, n = SHAPES[name]
            for m in [int(v) for v in args.ms.split(",")]:
                try:
                    r = time_shape(k, n, m, variant, args.graph, peak)
                except AssertionError as e:
                    r = {"error": str(e)}
                print(json.dumps({"variant": variant, "shape": name, "m": m, "graph": args.graph, **r}), flush=True)


if __name__ == "__main__":
    main()
