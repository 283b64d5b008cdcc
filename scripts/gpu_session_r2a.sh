#!/bin/bash
# round 2, session A: fixtures + parity numbers + reference column of config 3
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_contract.jsonl
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python tests/golden/make_tcnn_loss_curve.py seeds > gpurun_out/loss_gen_seeds.log 2>&1; echo "tcnn seeds rc=$?"
timeout 900 python scripts/loss_curves_ours.py > gpurun_out/loss_ours_seeds.log 2>&1; echo "ours seeds rc=$?"
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error|assert" gpurun_out/pytest_gpu.log | tail -30
timeout 1200 python scripts/microbench_tcnn.py > gpurun_out/microbench_tcnn.jsonl 2> gpurun_out/microbench_tcnn.err; echo "tcnn c3 rc=$?"
echo done
