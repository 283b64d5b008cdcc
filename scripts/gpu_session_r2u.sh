#!/bin/bash
# 1-GPU: full GPU suite (incl. the 128-neuron network with the warp-specialised inference kernel), wide microbench, N=1 bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python scripts/microbench_wide.py > gpurun_out/microbench_wide.jsonl 2> gpurun_out/microbench_wide.err; echo "microbench rc=$?"; cut -c1-330 gpurun_out/microbench_wide.jsonl; tail -3 gpurun_out/microbench_wide.err
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    j = json.loads(open('gpurun_out/bench_ours.json').read().strip().splitlines()[-1])
    print('N1', round(j['value']/1e9, 4), 'ms', round(j['ms_per_step'], 4), 'e2e ms', round(j['e2e']['ms_per_step'], 4), 'kernel ms', round(j['roofline']['ms_per_launch'], 4))
    for k, v in j.get('frame', {}).items():
        if isinstance(v, dict) and 'ms' in v: print(k, v['ms'])
        elif isinstance(v, dict):
            for k2, v2 in v.items():
                if isinstance(v2, dict) and 'ms' in v2: print(k, k2, v2['ms'])
except Exception as e: print('bench parse failed', e)
PY
echo done
