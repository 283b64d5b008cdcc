#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --no-frame > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "ours rc=$?"; tail -3 gpurun_out/bench_ours.err
python - <<'PY'
import json
j = json.load(open('gpurun_out/bench_ours.json'))
print({k: j[k] for k in ('value','ms_per_step','gpu_launches')}, j['e2e'], j['roofline']['ms_per_launch'], j['roofline']['frac'])
PY
timeout 600 ncu --set full --cache-control none --clock-control none --import-source on -k regex:"nrc_train_fused|nrc_adam|nrc_grid_ema" -s 30 -c 3 -o gpurun_out/prof_train python scripts/ncu_train_small.py > gpurun_out/ncu_train.log 2>&1; echo "ncu rc=$?"
echo done
