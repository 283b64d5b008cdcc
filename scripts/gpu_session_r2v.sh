#!/bin/bash
# 8-GPU box: scaling bench at N = 8 and N = 4 exactly as the driver launches it (incl. the 4K tile leg and the exchange check in the line)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 2951$1 bench.py --gpus $1 --steps 100 --warmup 10 2>gpurun_out/bench_n$1.err > gpurun_out/bench_n$1.json; echo "N=$1 rc=$?"
  python - <<PY
import json
try:
    j = json.loads(open('gpurun_out/bench_n$1.json').read().strip().splitlines()[-1]); g = j.get('gradient_exchange') or {}
    print('N', j['n_gpus'], 'value', round(j['value']/1e9, 4), 'ms', round(j['ms_per_step'], 4), 'e2e ms', round(j['e2e']['ms_per_step'], 4), 'floor', round(j['e2e'].get('transfer_floor_ms') or 0, 3),
          'solo', round(g.get('ms_per_step_single_gpu_schedule_without_exchange') or 0, 4), 'check', (g.get('check') or {}).get('ok'), 'tiles', (j.get('frame_4k_tiles') or {}))
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/bench_n$1.err').read()[-1500:])
PY
}
run 8
run 4
NRCHPM_OVERLAP=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 60 --warmup 10 --no-frame 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('N8 serial schedule: ms', round(j['ms_per_step'],4), 'e2e', round(j['e2e']['ms_per_step'],4))"
echo done
