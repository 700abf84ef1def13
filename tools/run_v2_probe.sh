#!/bin/bash
# one iteration of the fused-stage work loop on a B200: decoder parity tests, in-kernel stage clocks (needs the _trace side
# build: ONEBIT_LIB_SUFFIX=_trace ONEBIT_NVCC_EXTRA=-DONEBIT_TRACE python -m onebit_b200.build), bench line
python -m pytest tests/test_decoder_gpu.py -m gpu -x -q 2>&1 | tail -5
for st in 1 3 4 5; do ONEBIT_LIB_SUFFIX=_trace ONEBIT_TRACE_STAGE=$st python tools/trace_gemv2.py 2>&1 | tail -2; done
python bench.py --steps 32 --warmup 4 > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err; tail -c 300 gpurun_out/bench_v2.err
python -c "
import json;d=json.load(open('gpurun_out/bench_v2.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'],d['roofline']['us_per_launch'])"
