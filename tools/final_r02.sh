#!/bin/bash
# round-2 evidence run (one B200): smoke, GPU tests, bench line, module loop, launch list of a decode step,
# ncu --set full of the fused stages. Outputs under gpurun_out/ (copied into profiles/ by hand).
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/pytest_gpu_r02.log; cat gpurun_out/pytest_gpu_r02.log
python bench.py > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err; tail -c 400 gpurun_out/bench_r02_final.err
python tools/bench_module_loop.py > gpurun_out/module_loop.json 2> gpurun_out/module_loop.err; cat gpurun_out/module_loop.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fused_gemv2|attn_kernel|glue_kernel|lm_head|argmax|copy_ids" -s 700 -c 340 --csv --log-file gpurun_out/launches_b1_v2.csv python bench.py --steps 4 --warmup 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_gemv2_kernel -s 512 -c 4 -o gpurun_out/fused2_final_r02 -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_fused2_final.log 2>&1
tail -2 gpurun_out/ncu_fused2_final.log | cut -c1-200
