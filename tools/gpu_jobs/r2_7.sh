set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest7.log 2>&1
tail -3 gpurun_out/r2_pytest7.log
timeout 300 python tools/trace_step.py > gpurun_out/r2_trace5.log 2>&1
grep -n "opt_tail\|mb_begin\|minibatch duration\|main stream" gpurun_out/r2_trace5.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_bench5.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'])"
timeout 300 python bench.py --steps 3 --warmup 3 --batch 8192 --T 8192 --no-cpu-baseline > gpurun_out/r2_bench5_B8192.json 2> gpurun_out/r2_bench5b.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_bench5_B8192.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'])"
