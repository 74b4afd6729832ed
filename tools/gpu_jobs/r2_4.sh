set -x
mkdir -p gpurun_out
./tools/ubench/mma_rate > gpurun_out/r2_mma_rate.txt 2>&1
cat gpurun_out/r2_mma_rate.txt
timeout 900 python -m pytest tests/test_gpu_tc.py -q -x > gpurun_out/r2_pytest4.log 2>&1
tail -3 gpurun_out/r2_pytest4.log
timeout 900 python -m pytest tests/test_gpu_tc_epoch.py -q -s > gpurun_out/r2_parity4.log 2>&1
tail -8 gpurun_out/r2_parity4.log
timeout 300 python tools/trace_step.py > gpurun_out/r2_trace3.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
tail -c 300 gpurun_out/r2_bench3.json
