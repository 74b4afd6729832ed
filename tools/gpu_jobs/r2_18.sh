set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -q -x -k "mlp_chain" > gpurun_out/r2_pytest18a.log 2>&1
tail -12 gpurun_out/r2_pytest18a.log
timeout 1200 python -m pytest tests/test_gpu_tc.py tests/test_gpu_tc_epoch.py tests/test_gpu_collect.py -q > gpurun_out/r2_pytest18.log 2>&1
tail -12 gpurun_out/r2_pytest18.log
timeout 300 python tools/trace_step.py > gpurun_out/r2_trace9.log 2>&1
grep -n "minibatch duration\|main stream\|mlp_chain" gpurun_out/r2_trace9.log | head -20
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-fp32-tier --no-sweep --no-roofline > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_bench9.json').read().strip().splitlines()[-1]);print(round(d['value']),d['ms_per_step'],round(d['e2e']['value']),d['gpu_launches'])"
