#!/bin/bash
# 2-GPU session: C++ overlapped data-parallel schedule
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parallel.py tests/test_gpu_nrc.py -x -q -m gpu > gpurun_out/pytest_par.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_par.log
for ov in 1 0; do
NRCHPM_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 --no-frame > gpurun_out/bench_n2_ov$ov.json 2> gpurun_out/bench_n2_ov$ov.err; echo "bench n2 overlap=$ov rc=$?"; tail -3 gpurun_out/bench_n2_ov$ov.err
python - <<PY
import json
for l in open('gpurun_out/bench_n2_ov$ov.json'):
    if l.startswith('{'):
        j = json.loads(l); print({k: j[k] for k in ('value','ms_per_step')}, j['e2e']['ms_per_step'], j['e2e'].get('transfer_floor_ms'), json.dumps(j.get('gradient_exchange'))[:900])
PY
done
timeout 600 python bench.py --steps 100 --no-frame 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('N=1', {k: j[k] for k in ('value','ms_per_step')}, j['e2e'], j['roofline']['ms_per_launch'], json.dumps(j['cpu_baseline'])[:600])"
echo done
