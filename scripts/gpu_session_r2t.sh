#!/bin/bash
# 1-GPU: 128-neuron network -- tcnn fixtures (reference's own tiny-cuda-nn, width=128), parity tests, microbench vs tcnn
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tests/golden/make_tcnn_golden.py hash_ob_d6_w128 tri_ob_d5_w128 2>&1 | grep -i "wrote\|fail\|error" | cut -c1-200
cp gpurun_out/tcnn_*_w128.npz tests/golden/ 2>/dev/null; ls -la tests/golden/*w128* 2>&1 | cut -c1-120
timeout 600 python -m pytest tests/test_gpu_nrc.py -q -k "128" > gpurun_out/pytest_wide.log 2>&1; echo "pytest wide rc=$?"; tail -40 gpurun_out/pytest_wide.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_nrc.py tests/test_gpu_loss_curve.py tests/test_gpu_fullsize.py -q -x -k "not 128" > gpurun_out/pytest_narrow.log 2>&1; echo "pytest narrow rc=$?"; tail -3 gpurun_out/pytest_narrow.log
timeout 600 python scripts/microbench_wide.py > gpurun_out/microbench_wide.jsonl 2> gpurun_out/microbench_wide.err; echo "microbench rc=$?"; cut -c1-420 gpurun_out/microbench_wide.jsonl; tail -3 gpurun_out/microbench_wide.err
echo done
