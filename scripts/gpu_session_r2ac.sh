#!/bin/bash
# 1-GPU: after the tracker's register / occupancy change -- GPU suite, gen_rays timing in both schedules, bench line, ncu launch list of the bench command, ncu capture of gen_rays
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for c in 2 4; do for m in 1 2; do timeout 120 python scripts/tune_wavefront.py $m $c 2>&1 | tail -1; done; done | tee gpurun_out/tune_tracker_final.jsonl
timeout 900 python bench.py > gpurun_out/bench_ours.json 2>gpurun_out/bench_ours.err; echo "ours rc=$?"
python - <<'PY'
import json
j = json.loads(open('gpurun_out/bench_ours.json').read().strip().splitlines()[-1])
print({k: j[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'clocks')}); print('e2e', j['e2e']['ms_per_step'])
for k, v in j['frame']['config2_scene0_1080p'].items(): print(k, v['ms'])
print(json.dumps(j['frame'])[-1500:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_r2.csv python bench.py --steps 3 --warmup 3 --no-frame > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hpm_gen_rays -s 3 -c 1 -f -o gpurun_out/prof_gen_rays_r2_final python scripts/ncu_frame.py > gpurun_out/ncu_gen_rays.log 2>&1; echo "ncu gen_rays rc=$?"
echo done
