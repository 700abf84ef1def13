#!/usr/bin/env python
"""Timeline of one CTA of the tcgen05 decode-tile kernel (side build: ONEBIT_LIB_SUFFIX=_tc5trace
ONEBIT_NVCC_EXTRA=-DONEBIT_TC5_TRACE python -m onebit_b200.build; run with ONEBIT_LIB_SUFFIX=_tc5trace)."""
import ctypes
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from onebit_b200 import _lib, bitlinear  # noqa: E402

lib = _lib.load()
M, K, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
x = torch.randn(M, K, device="cuda", dtype=torch.float16)
w = torch.randint(-128, 128, (N, K // 8), dtype=torch.int8, device="cuda")
g = torch.rand(N, device="cuda", dtype=torch.float16) + 0.5
h = (torch.rand(K, device="cuda", dtype=torch.float16) * 3 - 1.5)
for _ in range(3):
    t = bitlinear.bitlinear_matvec(x, w, g, h, variant="tc5")
torch.cuda.synchronize()
buf = np.zeros(512, dtype=np.uint64)
rc = lib.onebit_debug_tc5_trace(buf.ctypes.data_as(ctypes.c_void_p))
assert rc == 0, rc
t = buf.astype(np.int64)
t0 = t[0]
print(f"M={M} K={K} N={N}: setup {t[1]-t0} ns, accumulators done {t[2]-t0}, epilogue done {t[3]-t0}, CTA end {t[4]-t0}")
n = K // 64
for c in range(min(n, 12)):
    a = t[16 + 4 * c: 20 + 4 * c] - t0
    f0, f1 = t[256 + 2 * c] - t0, t[257 + 2 * c] - t0
    print(f"  chunk {c:2d}: stage free {a[0]:6d}  stores issued {f0:6d} (+{f0-a[0]:4d})  fence done {f1:6d} (+{f1-f0:4d})  arrived {a[1]:6d}  "
          f"mma inputs ready {a[2]:6d}  mma issued {a[3]:6d}")
