#!/bin/bash
# 2-GPU box: the tests that need two devices (peer-memory optimizer step vs NCCL sum / vs its three-kernel form), smoke's world-2 exchange assertion
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parallel.py -x -q -rs > gpurun_out/pytest_gpu_parallel_n2.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_parallel_n2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 10 2>gpurun_out/bench_n2.err > gpurun_out/bench_n2.json; echo "bench N=2 rc=$?"
python - <<'PY'
import json
j = json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1]); g = j.get('gradient_exchange') or {}
print('N', j['n_gpus'], 'value', round(j['value']/1e9, 4), 'ms', round(j['ms_per_step'], 4), 'e2e ms', round(j['e2e']['ms_per_step'], 4), 'clocks', j['clocks'], 'check', (g.get('check') or {}).get('ok'), 'tiles', j.get('frame_4k_tiles'))
PY
echo done
