"""Prefill (large M) timing of the tcgen05 variant: TFLOP/s of t = S @ (h*x) (2*M*N*K flops) and of the full forward
(+ LayerNorm), CUDA events, per LLaMA shape. Activations M x K fp16 (>= L2 for M >= 8192)."""
import argparse, json, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from onebit_b200 import _lib

SHAPES = {"7b_attn": (4096, 4096), "7b_gate_up": (4096, 11008), "7b_down": (11008, 4096), "13b_gate_up": (5120, 13824)}
ap = argparse.ArgumentParser()
ap.add_argument("--ms", default="512,2048,16384")
ap.add_argument("--shapes", default=",".join(SHAPES))
args = ap.parse_args()
lib = _lib.load()
dev = torch.device("cuda:0")
root = Path(__file__).resolve().parent.parent
peaks = json.loads((root / "MEASURED_PEAKS.json").read_text()) if (root / "MEASURED_PEAKS.json").exists() else {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
for name in args.shapes.split(","):
    k, n = SHAPES[name]
    w = torch.randint(-128, 128, (n, k // 8), dtype=torch.int8, device=dev)
    g = (torch.rand(n, device=dev) + 0.5).half()
    h = (torch.rand(k, device=dev) * 3 - 1.5).half()
    for m in [int(v) for v in args.ms.split(",")]:
        x = torch.randn(m, k, device=dev).half()
        t = torch.empty(m, n, dtype=torch.float32, device=dev)
        y = torch.empty(m, n, dtype=torch.float16, device=dev)
        wsb = lib.onebit_bitlinear_workspace_bytes(m, k, n)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        mws = lib.onebit_matvec_workspace_bytes(m, k)
        st = torch.cuda.current_stream().cuda_stream
        res = {"shape": name, "K": k, "N": n, "M": m}
        for which in ("matvec", "forward"):
            def run():
                if which == "matvec":
                    rc = lib.onebit_bitlinear_matvec(x.data_ptr(), w.data_ptr(), g.data_ptr(), h.data_ptr(), t.data_ptr(), m, k, n, 0, 0, 1, ws.data_ptr(), mws, _lib.VARIANT_TC5, st)
                else:
                    rc = lib.onebit_bitlinear_forward(x.data_ptr(), w.data_ptr(), g.data_ptr(), h.data_ptr(), None, y.data_ptr(), m, k, n, 0, 0, 1e-5, ws.data_ptr(), wsb, _lib.VARIANT_TC5, st)
                assert rc == 0, _lib.last_error()
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            e0.record()
            for _ in range(reps):
                run()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / reps
            tf = 2.0 * m * n * k / us / 1e6
            res[which] = {"us": round(us, 1), "TFLOPs": round(tf, 1), "frac_burst": round(tf / peaks["bf16_tflops"], 3),
                          "frac_sustained": round(tf / peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]), 3)}
        print(json.dumps(res), flush=True)
