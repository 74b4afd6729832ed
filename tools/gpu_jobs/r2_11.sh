set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err
tail -5 gpurun_out/r2_bench8.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_bench8.json').read().strip().splitlines()[-1])
print(round(d['value']),d['ms_per_step'],round(d['e2e']['value']))
for k in ('strong_sweep','fp32_tier','gae','cpu_baseline'): print(k, json.dumps(d.get(k))[:600])
r=d.get('roofline',{}); print('roofline', r.get('kernel'), r.get('frac'), r.get('us_per_minibatch'), r.get('error'))
for o in r.get('other_kernels',[]): print('  ', o['kernel'], round(o['frac'],4), round(o['us_per_minibatch'],1), o['launches_per_minibatch'])
"
timeout 900 ncu --set full --clock-control none -s 116 -c 114 -o gpurun_out/r2_full python tools/profile_step.py --minibatches 2 > gpurun_out/r2_ncu_full.log 2>&1
tail -2 gpurun_out/r2_ncu_full.log
