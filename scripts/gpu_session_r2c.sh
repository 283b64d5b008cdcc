#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python scripts/train_profile.py > gpurun_out/train_profile.jsonl 2> gpurun_out/train_profile.err; echo "profile rc=$?"; cat gpurun_out/train_profile.jsonl; tail -3 gpurun_out/train_profile.err
timeout 300 python scripts/tune_train.py > gpurun_out/tune_train.jsonl 2> gpurun_out/tune_train.err; echo "tune_train rc=$?"; cat gpurun_out/tune_train.jsonl | cut -c1-400; tail -3 gpurun_out/tune_train.err
timeout 900 python -m pytest tests/test_gpu_nrc.py -x -q -m gpu > gpurun_out/pytest_nrc.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_nrc.log
echo done
