set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -s 116 -c 114 -o /tmp/r2_full python tools/profile_step.py --minibatches 2 > gpurun_out/r2_ncu_full.log 2>&1
tail -2 gpurun_out/r2_ncu_full.log
ncu -i /tmp/r2_full.ncu-rep --page raw --csv > gpurun_out/r2_full_raw.csv 2>/dev/null
ls -la gpurun_out/r2_full_raw.csv /tmp/r2_full.ncu-rep
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err
tail -c 300 gpurun_out/r2_bench8.json
