set -x
mkdir -p gpurun_out
./tools/ubench/mma_rate > gpurun_out/r2_mma_rate2.txt 2>&1
head -12 gpurun_out/r2_mma_rate2.txt
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_tc_epoch.py -q -x > gpurun_out/r2_pytest6.log 2>&1
tail -3 gpurun_out/r2_pytest6.log
timeout 300 python tools/trace_step.py > gpurun_out/r2_trace4.log 2>&1
grep -n "opt_tail\|mb_begin\|minibatch duration\|main stream" gpurun_out/r2_trace4.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_bench4.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'])"
