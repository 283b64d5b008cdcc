#!/bin/bash
# 2-GPU session: overlapped schedule variants (head fraction)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 --steps 100 --warmup 10 --no-frame 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        j = json.loads(l); g = j.get('gradient_exchange') or {}
        print('$2', j['n_gpus'], {k: round(j[k], 4) for k in ('value','ms_per_step')}, 'e2e', round(j['e2e']['ms_per_step'], 3), 'floor', round(j['e2e'].get('transfer_floor_ms') or 0, 3), 'solo', round(g.get('ms_per_step_single_gpu_schedule_without_exchange') or 0, 4), (g.get('check') or {}).get('ok'))"
}
for head in 0.0 0.3 0.5; do NRCHPM_OVERLAP_HEAD=$head run 2 "head=$head"; done
NRCHPM_OVERLAP=0 run 2 "serial"
echo done
