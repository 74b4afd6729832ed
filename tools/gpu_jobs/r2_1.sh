set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_smi.txt 2>&1
nproc > gpurun_out/r2_cpu.txt; free -g >> gpurun_out/r2_cpu.txt
timeout 900 python -m pytest tests/test_gpu_tc_epoch.py -q -s > gpurun_out/r2_parity1.log 2>&1
timeout 300 python tools/trace_step.py > gpurun_out/r2_trace0.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_wgrad_kernel -s 44 -c 22 -o gpurun_out/r2_wgrad python tools/profile_step.py --minibatches 2 > gpurun_out/r2_ncu_wgrad.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench0.json 2> gpurun_out/r2_bench0.err
tail -3 gpurun_out/r2_parity1.log
