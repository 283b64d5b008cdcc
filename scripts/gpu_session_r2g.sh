#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python scripts/tune_infer_variants.py > gpurun_out/tune_infer_variants.jsonl 2>&1; echo "variants rc=$?"; cat gpurun_out/tune_infer_variants.jsonl | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_nrc.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/pytest_nrc.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_nrc.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nrc_infer_ws -s 3 -c 1 -o gpurun_out/prof_infer_ws python scripts/ncu_infer.py > gpurun_out/ncu_infer_ws.log 2>&1; echo "ncu rc=$?"
echo done
