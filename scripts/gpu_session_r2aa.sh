#!/bin/bash
# 1-GPU: tracker lookup diet (integer floor / range test, LUT base in a register): parity suite of the tracker + timing of the gen_rays pass
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tracker.py tests/test_gpu_configs.py tests/test_gpu_wavefront.py tests/test_gpu_frames.py tests/test_gpu_nrc.py -x -q > gpurun_out/pytest_tracker.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_tracker.log
for c in 2 4; do for m in 1 2; do timeout 120 python scripts/tune_wavefront.py $m $c 2>&1 | tail -1; done; done | tee gpurun_out/tune_tracker_diet.jsonl
echo done
