for st in 1 3 5; do ONEBIT_DUP_STAGE=1 ONEBIT_LIB_SUFFIX=_trace ONEBIT_TRACE_STAGE=$st python tools/trace_gemv2.py 2>&1 | tail -2; done
ONEBIT_DUP_STAGE=1 python bench.py --steps 32 --warmup 4 > gpurun_out/bench_v2_dup.json 2> gpurun_out/bench_v2_dup.err; tail -c 300 gpurun_out/bench_v2_dup.err
python -c "
import json;d=json.load(open('gpurun_out/bench_v2_dup.json'));print('dup', d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'],d['roofline']['us_per_launch'])"
