python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/pytest_gpu_r02.log; cat gpurun_out/pytest_gpu_r02.log
python bench.py > gpurun_out/bench_r02_final3.json 2> gpurun_out/bench_r02_final3.err; tail -c 300 gpurun_out/bench_r02_final3.err
python -c "
import json;d=json.load(open('gpurun_out/bench_r02_final3.json'));b=d['batch32'];print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'],'| b32',b['value'],b['ms_per_step'],b['launches_per_step'],b['roofline']['us_per_launch'],b['roofline']['frac'],'| c3',d['prefill2048_decode128_b8']['decode_ms_per_step'])"
