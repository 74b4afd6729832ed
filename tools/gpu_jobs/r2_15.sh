set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest15.log 2>&1
tail -40 gpurun_out/r2_pytest15.log
timeout 300 python tools/trace_step.py > gpurun_out/r2_trace8.log 2>&1
grep -n "minibatch duration\|main stream" gpurun_out/r2_trace8.log
