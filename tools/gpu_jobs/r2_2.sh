set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc_epoch.py -q -s -x > gpurun_out/r2_parity2.log 2>&1
tail -3 gpurun_out/r2_parity2.log
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_kernels.py -q -x > gpurun_out/r2_pytest2.log 2>&1
tail -3 gpurun_out/r2_pytest2.log
timeout 600 python -m pytest tests/test_gpu_ppo.py -q -x > gpurun_out/r2_pytest2b.log 2>&1
tail -3 gpurun_out/r2_pytest2b.log
timeout 300 python tools/trace_step.py > gpurun_out/r2_trace1.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
tail -c 600 gpurun_out/r2_bench1.json
