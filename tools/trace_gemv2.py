"""Stage clocks (clock64 of CTA 0) of one second-generation fused stage inside a real decode step. Needs the _trace side
build: ONEBIT_LIB_SUFFIX=_trace ONEBIT_NVCC_EXTRA=-DONEBIT_TRACE python -m onebit_b200.build; run with
ONEBIT_LIB_SUFFIX=_trace ONEBIT_TRACE_STAGE=<1|3|4|5> (qkv | o | gate,up | down)."""
import ctypes, os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from onebit_b200 import BitLlamaDecoderB200, LLAMA_7B, synthetic_state_dict, _lib

cfg = dict(LLAMA_7B, num_hidden_layers=4)
dec = BitLlamaDecoderB200(cfg, synthetic_state_dict(cfg), max_seq_len=128)
dec.reset(torch.tensor([5]))
for _ in range(20):
    dec.step()
torch.cuda.synchronize()
lib = _lib.load()
out = (ctypes.c_longlong * 16)()
names = "wait-done | records | elementwise | quantise | barrier+weights | IMMA | epilogue"
print("stage", os.environ.get("ONEBIT_TRACE_STAGE", "5"), names)
for rep in range(4):
    dec.step()
    torch.cuda.synchronize()
    assert lib.onebit_debug_read_trace16(out) == 0
    t = list(out)
    print([t[i + 1] - t[i] for i in range(7)], "total", t[7] - t[0],
          "| first load back +", t[13] - t[1], "args in smem +", t[14] - t[0], "signs issued +", t[15] - t[0], "pre-wait", t[12] - t[0], "wait", t[1] - t[12], "loads issued +", t[8] - t[1], "records in +", t[9] - t[1], "butterfly +", t[10] - t[1],
          "scalars +", t[11] - t[1], "barrier +", t[2] - t[1])
