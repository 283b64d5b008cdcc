#!/bin/bash
# full GPU suite + both bench arms
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "ours rc=$?"; tail -3 gpurun_out/bench_ours.err
python - <<'PY'
import json
j = json.load(open('gpurun_out/bench_ours.json'))
print({k: j[k] for k in ('value','ms_per_step','gpu_launches')}, j['e2e'], j['roofline']['ms_per_launch'], j['roofline']['frac'])
print(json.dumps(j.get('frame', {}))[:1500])
PY
echo done
