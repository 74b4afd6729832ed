mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --no-fp32-tier --no-cpu-baseline > gpurun_out/r2_final_bench_2gpu.json 2> gpurun_out/r2_final_bench_2gpu.err
tail -3 gpurun_out/r2_final_bench_2gpu.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_final_bench_2gpu.json').read().strip().splitlines()[-1])
print(round(d['value']),d['ms_per_step'],round(d['e2e']['value'])); print(json.dumps(d.get('strong_sweep'))[:400])"
