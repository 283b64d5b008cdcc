#!/bin/bash
# round 2, session B: fused training kernel + split optimizer
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python scripts/tune_train.py > gpurun_out/tune_train.jsonl 2> gpurun_out/tune_train.err; echo "tune_train rc=$?"; cat gpurun_out/tune_train.jsonl | cut -c1-400; tail -3 gpurun_out/tune_train.err
timeout 900 python -m pytest tests/test_gpu_nrc.py tests/test_gpu_loss_curve.py tests/test_gpu_fullsize.py tests/test_gpu_configs.py -x -q -m gpu > gpurun_out/pytest_nrc.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_nrc.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/launches_train.csv python scripts/ncu_train_small.py > gpurun_out/ncu_train.log 2>&1; echo "ncu rc=$?"
echo done
