#!/bin/bash
# final round-2 validation: the whole GPU suite, per-size kernel table, the bench line, smoke, obs-kernel timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_final_tests.log 2>&1
echo "suite exit $?" >> gpurun_out/r2_final_tests.log
timeout 120 python -m tools.bench_obs > gpurun_out/r2_obs_kernels.txt 2>&1
timeout 300 python -m tools.trace_sizes --batches 16384 32768 65536 > gpurun_out/r2_trace_sizes_after.txt 2>&1
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
timeout 120 python __graft_entry__.py --smoke > gpurun_out/r2_final_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/r2_final_smoke.log
tail -n 4 gpurun_out/r2_final_tests.log; tail -n 3 gpurun_out/r2_final_smoke.log; grep -v Warn gpurun_out/r2_trace_sizes_after.txt; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_final_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d.get("strong_sweep", {}).get("value"), d.get("fp32_tier", {}).get("value"))
PY
