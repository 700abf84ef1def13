python -m pytest tests/test_decoder_gpu.py -m gpu -x -q 2>&1 | tail -5
for th in ${THS:-512}; do
for st in 1 4 5; do ONEBIT_FUSED2_THREADS=$th ONEBIT_LIB_SUFFIX=_trace ONEBIT_TRACE_STAGE=$st python tools/trace_gemv2.py 2>&1 | tail -2; done
ONEBIT_FUSED2_THREADS=$th python bench.py --steps 32 --warmup 4 > gpurun_out/bench_v2_$th.json 2> gpurun_out/bench_v2_$th.err; tail -c 300 gpurun_out/bench_v2_$th.err
python -c "
import json;d=json.load(open('gpurun_out/bench_v2_$th.json'));print($th, d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'],d['roofline']['us_per_launch'])"
done
