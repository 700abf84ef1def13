"""Stage timing (clock64 of CTA 0) of the last GEMV launch inside a real decode step. Needs the _trace side build:
ONEBIT_LIB_SUFFIX=_trace ONEBIT_NVCC_EXTRA=-DONEBIT_TRACE python -m onebit_b200.build; run with ONEBIT_LIB_SUFFIX=_trace."""
import ctypes, sys
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
out = (ctypes.c_longlong * 8)()
names = ["start->weights issued", "->pdl wait entered", "->wait done", "->digits in smem", "->IMMA done", "->end"]
for rep in range(5):
    dec.step()
    torch.cuda.synchronize()
    import os
    fn = lib.onebit_debug_read_trace_decoder if os.environ.get('ONEBIT_FUSED', '1') != '0' else lib.onebit_debug_read_trace
    assert fn(out) == 0
    t = list(out)[:8]
    print("last GEMV-stage of the step (down_proj), stage cycles:", [t[i + 1] - t[i] for i in range(7)], "total", t[7] - t[0])
g = torch.cuda.CUDAGraph()
