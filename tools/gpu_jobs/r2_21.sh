set -x
mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=1 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-fp32-tier --no-roofline --no-graph --sweep-envs 1 --sweep-batch 8192 --sweep-transitions 131072 > gpurun_out/r2_sweep_e1.json 2> gpurun_out/r2_sweep_e1.err
tail -c 1500 gpurun_out/r2_sweep_e1.json
grep -v "^$" gpurun_out/r2_sweep_e1.err | tail -30
