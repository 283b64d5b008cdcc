#!/bin/bash
# 1-GPU: the round-end sequence as the driver runs it -- GPU suite, smoke, reference arm, our arm (default flags)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2>gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_ref.json
timeout 900 python bench.py > gpurun_out/bench_ours.json 2>gpurun_out/bench_ours.err; echo "ours rc=$?"
python - <<'PY'
import json
j = json.loads(open('gpurun_out/bench_ours.json').read().strip().splitlines()[-1])
print({k: j[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'clocks')}); print('e2e', j['e2e']); print('roofline', {k: j['roofline'][k] for k in ('achieved', 'frac', 'ms_per_launch', 'traffic')})
print('frame', json.dumps(j.get('frame'))[:1500])
PY
echo done
