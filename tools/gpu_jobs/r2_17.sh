set -x
mkdir -p gpurun_out
for sb in 16384 32768 65536 131072; do
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-fp32-tier --no-roofline --sweep-batch $sb --sweep-transitions 524288 > gpurun_out/r2_sweep_$sb.json 2> gpurun_out/r2_sweep_$sb.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_sweep_$sb.json').read().strip().splitlines()[-1]);s=d['strong_sweep'];print($sb, round(s.get('value',0)), s.get('ms_per_step'), s.get('error'), s.get('skipped'))"
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv
