#!/bin/bash
# 1-GPU: round-end sequence on the final library -- C-ABI tests on a GPU box, GPU suite, smoke, both bench arms
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cabi.py -x -q 2>&1 | tail -2
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2>gpurun_out/bench_ref.err; echo "ref rc=$?"; python -c "
import json; j=json.loads(open('gpurun_out/bench_ref.json').read().strip().splitlines()[-1]); print('ref', j['value'], j['ms_per_step'], j['clocks'])"
timeout 900 python bench.py > gpurun_out/bench_ours.json 2>gpurun_out/bench_ours.err; echo "ours rc=$?"
python - <<'PY'
import json
j = json.loads(open('gpurun_out/bench_ours.json').read().strip().splitlines()[-1])
print({k: j[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'clocks')}); print('e2e', j['e2e']['ms_per_step'], 'roofline frac', j['roofline']['frac'], j['roofline']['ms_per_launch'])
for k, v in j['frame']['config2_scene0_1080p'].items(): print(k, v['ms'])
PY
echo done
