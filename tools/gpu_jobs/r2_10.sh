set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_tc_epoch.py tests/test_gpu_kernels.py -q -x > gpurun_out/r2_pytest10.log 2>&1
tail -2 gpurun_out/r2_pytest10.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench7.json 2> gpurun_out/r2_bench7.err
timeout 300 python bench.py --steps 3 --warmup 3 --batch 8192 --T 8192 --no-cpu-baseline > gpurun_out/r2_bench7_B8192.json 2>> gpurun_out/r2_bench7.err
timeout 300 python bench.py --steps 2 --warmup 2 --batch 65536 --T 8192 --no-cpu-baseline > gpurun_out/r2_bench7_B65536.json 2>> gpurun_out/r2_bench7.err
for f in gpurun_out/r2_bench7.json gpurun_out/r2_bench7_B8192.json gpurun_out/r2_bench7_B65536.json; do python -c "
import json,sys;d=json.loads(open('$f').read().strip().splitlines()[-1]);print('$f',round(d['value']),d['ms_per_step'],round(d['e2e']['value']))"; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 114 -c 230 --csv --log-file gpurun_out/r2_f16_launches.csv python tools/profile_step.py --minibatches 3 > gpurun_out/r2_ncu_launches.log 2>&1
timeout 300 python tools/trace_step.py > gpurun_out/r2_trace7.log 2>&1
grep -n "opt_\|mb_begin\|loss\|minibatch duration\|main stream" gpurun_out/r2_trace7.log
