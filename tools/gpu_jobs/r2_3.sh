set -x
mkdir -p gpurun_out
timeout 300 python tools/trace_step.py > gpurun_out/r2_trace2.log 2>&1
timeout 600 python -m pytest tests/test_gpu_tc_epoch.py -q -s -x -k "step_all_keys" > gpurun_out/r2_parity3.log 2>&1
tail -3 gpurun_out/r2_parity3.log
timeout 900 python -m pytest tests/test_gpu_tc.py -q -x > gpurun_out/r2_pytest3.log 2>&1
tail -3 gpurun_out/r2_pytest3.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
tail -c 300 gpurun_out/r2_bench2.json
