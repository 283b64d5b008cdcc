#!/bin/bash
# 2-GPU session: sharded exchange
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/check_peer_exchange.py 2>&1 | grep "^{" ; echo "check rc=${PIPESTATUS[0]}"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/overlap_timeline.py 2>&1 | grep -v "^\*\|OMP_NUM\|NCCL\|^$" | head -14
run() {
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 --steps 100 --warmup 10 --no-frame 2>gpurun_out/bench_err.log | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        j = json.loads(l); g = j.get('gradient_exchange') or {}
        print('$2', j['n_gpus'], {k: round(j[k], 4) for k in ('value','ms_per_step')}, 'e2e', round(j['e2e']['ms_per_step'], 3), 'floor', round(j['e2e'].get('transfer_floor_ms') or 0, 3), 'solo', round(g.get('ms_per_step_single_gpu_schedule_without_exchange') or 0, 4), (g.get('check') or {}))"
  tail -2 gpurun_out/bench_err.log | cut -c1-300
}
run 2 "fused-serial"; NRCHPM_PEER_FUSED=0 NRCHPM_OVERLAP=1 run 2 "steps-overlap"
NRCHPM_PEER_FUSED=0 run 2 "steps-serial"
echo done
