set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest9.log 2>&1
tail -3 gpurun_out/r2_pytest9.log
timeout 300 python tools/trace_step.py > gpurun_out/r2_trace6.log 2>&1
grep -n "opt_\|mb_begin\|loss\|minibatch duration\|main stream" gpurun_out/r2_trace6.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_bench6.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'])"
