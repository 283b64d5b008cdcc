#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for p in 0 1; do
  echo "== NRCHPM_L2_PERSIST=$p"
  NRCHPM_L2_PERSIST=$p timeout 300 python scripts/train_profile.py 2 2>&1 | cut -c1-1200
  NRCHPM_L2_PERSIST=$p timeout 300 python - <<'PY'
import sys, json
sys.path.insert(0, 'scripts'); sys.path.insert(0, '.')
import tune_train
print(json.dumps(tune_train.run(1, 2))[:400])
PY
  NRCHPM_L2_PERSIST=$p timeout 600 python bench.py --no-frame --steps 200 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print({k: j[k] for k in ('value','ms_per_step')}, j['e2e']['ms_per_step'], j['roofline']['ms_per_launch'])"
done
echo done
