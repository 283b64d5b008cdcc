#!/bin/bash
# 2-GPU box: full GPU suite (incl. the world-2 peer exchange test), N=1 bench, N=2 serial vs overlapped schedule (cp.async peer kernel)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    j = json.loads(open('gpurun_out/bench_ours.json').read().strip().splitlines()[-1])
    print('N1', round(j['value']/1e9, 4), 'ms', round(j['ms_per_step'], 4), 'e2e ms', round(j['e2e']['ms_per_step'], 4), 'kernel ms', round(j['roofline']['ms_per_launch'], 4))
    f = j.get('frame', {})
    for k, v in f.items():
        if isinstance(v, dict) and 'ms' in v: print(k, v['ms'])
        elif isinstance(v, dict):
            for k2, v2 in v.items():
                if isinstance(v2, dict) and 'ms' in v2: print(k, k2, v2['ms'])
except Exception as e: print('bench parse failed', e)
PY
run() {
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 60 --warmup 10 --no-frame 2>gpurun_out/bench_err.log | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        open('gpurun_out/bench_n2_$2.json','w').write(l)
        j = json.loads(l); g = j.get('gradient_exchange') or {}
        print('$1', j['n_gpus'], {k: round(j[k], 4) for k in ('value','ms_per_step')}, 'e2e', round(j['e2e']['ms_per_step'], 3), 'solo', round(g.get('ms_per_step_single_gpu_schedule_without_exchange') or 0, 4), (g.get('check') or {}).get('ok'))"
  tail -2 gpurun_out/bench_err.log | grep -v "OMP_NUM\|^\*\*\*" | cut -c1-300
}
NRCHPM_OVERLAP=0 run "serial" serial
for sms in 96 112 128; do NRCHPM_OVERLAP=1 NRCHPM_OVERLAP_SMS=$sms run "overlap sms=$sms" ov$sms; done
NRCHPM_OVERLAP=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/overlap_timeline.py 2>&1 | grep -v "^\*\|OMP_NUM\|NCCL\|^$" | head -20
echo done
