#!/bin/bash
# 2-GPU box: 4K frame on tiles with the data-parallel training pipelined under the next frame's tracking
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 60 --warmup 10 2>gpurun_out/bench_n$N.err > gpurun_out/bench_n$N.json; echo "bench N=$N rc=$?"
python - <<PY
import json
try:
    j = json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
    print('N', j['n_gpus'], 'ms', round(j['ms_per_step'], 4), 'check', ((j.get('gradient_exchange') or {}).get('check') or {}).get('ok'), 'tiles', j.get('frame_4k_tiles'))
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/bench_n$N.err').read()[-2000:])
PY
echo done
