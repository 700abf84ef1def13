python -m pytest tests/test_bitlinear_gpu.py -m gpu -q -x 2>&1 | tail -4
python tools/bench_module_loop.py > gpurun_out/module_loop.json 2> gpurun_out/module_loop.err; cat gpurun_out/module_loop.json; tail -3 gpurun_out/module_loop.err
python bench.py > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err; tail -c 300 gpurun_out/bench_r02_final.err
python -c "
import json;d=json.load(open('gpurun_out/bench_r02_final.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'],d['roofline']['us_per_launch'],d['roofline']['traffic'], d.get('batch32',{}).get('value'), d.get('batch32',{}).get('ms_per_step'))"
