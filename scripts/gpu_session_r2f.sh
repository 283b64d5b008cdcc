#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python scripts/tune_infer.py > gpurun_out/tune_infer.jsonl 2> gpurun_out/tune_infer.err; echo "tune_infer rc=$?"; cat gpurun_out/tune_infer.jsonl | cut -c1-500; tail -3 gpurun_out/tune_infer.err
timeout 900 python -m pytest tests/test_gpu_nrc.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/pytest_nrc.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_nrc.log
echo done
