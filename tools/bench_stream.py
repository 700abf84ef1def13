"""Streaming limit of the IMMA GEMV kernel: ONE launch over a tall stacked sign matrix (>= 3x L2), M = 1..2.
This is the kernel timed alone on inputs larger than L2, free of per-launch latency."""
import json, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from onebit_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda:0")
peak = json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] \
    if (Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").exists() else 6650.0
for k in (4096, 11008):
    for m in (1, 2):
        n = int(400e6 // (k // 8)) // 32 * 32
        w = torch.randint(-128, 128, (n, k // 8), dtype=torch.int8, device=dev)
        g = (torch.rand(n, device=dev) + 0.5).half()
        h = (torch.rand(k, device=dev) * 3 - 1.5).half()
        x = torch.randn(m, k, device=dev).half()
        t = torch.empty(m, n, dtype=torch.float32, device=dev)
        wsb = lib.onebit_matvec_workspace_bytes(m, k)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        def run():
            rc = lib.onebit_bitlinear_matvec(x.data_ptr(), w.data_ptr(), g.data_ptr(), h.data_ptr(), t.data_ptr(), m, k, n,
                                             0, 0, 1, ws.data_ptr(), wsb, _lib.VARIANT_MMA, st)
            assert rc == 0, _lib.last_error()
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        by = n * k // 8 + m * (k // 256) * 1024 + 4 * m * n + 2 * n
        print(json.dumps({"K": k, "N": n, "M": m, "us": round(us, 1), "GBs": round(by / us / 1e3, 1),
                          "frac_of_measured_hbm": round(by / us / 1e3 / peak, 3)}), flush=True)
