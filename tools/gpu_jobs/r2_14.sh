set -x
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_dp2.py -q -s -x > gpurun_out/r2_dp2.log 2>&1
tail -30 gpurun_out/r2_dp2.log
