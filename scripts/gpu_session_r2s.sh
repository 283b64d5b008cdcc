#!/bin/bash
# 1-GPU: EMA pass inside the training launch + programmatic dependent launch of the training chain
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python scripts/l2_probe.py
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for cfg in "1 1" "0 1" "1 0" "0 0"; do set -- $cfg
  echo "== EMA_IN_KERNEL=$1 PDL=$2"
  NRCHPM_EMA_IN_KERNEL=$1 NRCHPM_PDL=$2 timeout 300 python scripts/train_profile.py 2 2>&1 | tail -1 | cut -c1-1400
  NRCHPM_EMA_IN_KERNEL=$1 NRCHPM_PDL=$2 timeout 300 python bench.py --steps 200 --warmup 10 --no-frame 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('bench ms_per_step', round(j['ms_per_step'],4), 'kernel', round(j['roofline']['ms_per_launch'],4), 'e2e', round(j['e2e']['ms_per_step'],4), 'loss', j['loss'])"
done
echo done
