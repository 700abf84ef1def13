for st in 1 3 5; do ONEBIT_LIB_SUFFIX=_trace ONEBIT_TRACE_STAGE=$st python tools/trace_gemv2.py 2>&1 | tail -2; done
ncu --set full --clock-control none --import-source on -k regex:fused_gemv2_kernel -s 200 -c 4 -o gpurun_out/fused2_r02 -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_fused2.log 2>&1
tail -3 gpurun_out/ncu_fused2.log
