set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest19.log 2>&1
tail -6 gpurun_out/r2_pytest19.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench10.json 2> gpurun_out/r2_bench10.err
tail -3 gpurun_out/r2_bench10.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_bench10.json').read().strip().splitlines()[-1])
print(round(d['value']),d['ms_per_step'],round(d['e2e']['value']),d['gpu_launches'])
print('sweep', d['strong_sweep'].get('value'), 'fp32', d['fp32_tier'].get('value'), 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['as_shipped']['value'])
r=d['roofline']; print(r['kernel'], r['frac'], r['traffic'], r.get('ncu'))"
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2_smoke.log 2>&1
tail -3 gpurun_out/r2_smoke.log
