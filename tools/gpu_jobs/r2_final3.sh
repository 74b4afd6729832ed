#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_obs.py tests/test_gpu_tc.py tests/test_gpu_tc_epoch.py -q > gpurun_out/r2_final3_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2_final3_tests.log
timeout 400 python -m tools.trace_sizes --batches 16384 32768 65536 > gpurun_out/r2_trace_sizes_after.txt 2>&1
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_final3_bench.json 2> gpurun_out/r2_final3_bench.err
tail -n 4 gpurun_out/r2_final3_tests.log; grep -v Warn gpurun_out/r2_trace_sizes_after.txt; head -c 600 gpurun_out/r2_final3_bench.json
