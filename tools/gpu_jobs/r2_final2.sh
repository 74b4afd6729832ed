#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_obs.py tests/test_gpu_a2c.py -q -s > gpurun_out/r2_final_new.log 2>&1
echo "new tests exit $?" >> gpurun_out/r2_final_new.log
timeout 120 python -m tools.bench_obs > gpurun_out/r2_obs_kernels.txt 2>&1
timeout 400 python -m tools.trace_sizes --batches 16384 32768 65536 > gpurun_out/r2_trace_sizes.txt 2>&1
tail -n 8 gpurun_out/r2_final_new.log; cat gpurun_out/r2_obs_kernels.txt; cat gpurun_out/r2_trace_sizes.txt
