#!/bin/bash
# 1-GPU: analytic sky cull in gen_rays (bit-exactness + time), ncu source-level capture of gen_rays, ncu capture of the shipped inference kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tracker.py tests/test_gpu_frames.py tests/test_gpu_configs.py -x -q > gpurun_out/pytest_tracker.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_tracker.log
for s in 0 5; do timeout 120 python scripts/ncu_frame.py $s; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hpm_gen_rays -s 3 -c 1 -f -o gpurun_out/prof_gen_rays_cull python scripts/ncu_frame.py > gpurun_out/ncu_gen_rays.log 2>&1; echo "ncu gen_rays rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nrc_infer_ws -s 3 -c 1 -f -o gpurun_out/prof_infer_ws_shipped python scripts/ncu_infer.py > gpurun_out/ncu_infer_ws.log 2>&1; echo "ncu infer rc=$?"
echo done
