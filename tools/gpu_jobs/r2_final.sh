#!/bin/bash
# final round-2 validation: new-row tests first, the whole GPU suite, the bench line, smoke, obs-kernel timings
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_obs.py tests/test_gpu_a2c.py -q -s -x > gpurun_out/r2_final_new.log 2>&1
echo "new tests exit $?" >> gpurun_out/r2_final_new.log
timeout 120 python -m tools.bench_obs > gpurun_out/r2_obs_kernels.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_final_tests.log 2>&1
echo "suite exit $?" >> gpurun_out/r2_final_tests.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
timeout 120 python __graft_entry__.py --smoke > gpurun_out/r2_final_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/r2_final_smoke.log
tail -3 gpurun_out/r2_final_new.log gpurun_out/r2_final_tests.log gpurun_out/r2_final_smoke.log; cat gpurun_out/r2_obs_kernels.txt; head -c 1500 gpurun_out/r2_final_bench.json
