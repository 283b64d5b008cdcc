#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
run() {
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 60 --warmup 10 --no-frame 2>gpurun_out/bench_err.log | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        j = json.loads(l); g = j.get('gradient_exchange') or {}
        print('$1', j['n_gpus'], {k: round(j[k], 4) for k in ('value','ms_per_step')}, 'solo', round(g.get('ms_per_step_single_gpu_schedule_without_exchange') or 0, 4), (g.get('check') or {}).get('ok'))"
  tail -2 gpurun_out/bench_err.log | grep -v "OMP_NUM\|^\*\*\*" | cut -c1-300
}
run "serial"
for sms in 108; do for head in 0.0; do NRCHPM_OVERLAP=1 NRCHPM_OVERLAP_SMS=$sms NRCHPM_OVERLAP_HEAD=$head run "overlap sms=$sms head=$head"; done; done
echo done
