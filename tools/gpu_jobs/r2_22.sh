set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_collect.py -q > gpurun_out/r2_pytest22.log 2>&1
tail -4 gpurun_out/r2_pytest22.log
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-fp32-tier --no-roofline --sweep-envs 1 --sweep-batch 8192 --sweep-transitions 131072 > gpurun_out/r2_sweep_e1.json 2> gpurun_out/r2_sweep_e1.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_sweep_e1.json').read().strip().splitlines()[-1]);print(json.dumps(d['strong_sweep'])[:300])"
