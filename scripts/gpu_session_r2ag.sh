#!/bin/bash
# 1-GPU: last EMA pass of nrc_infer_and_train_host launched behind the inference kernels: host-path parity tests, e2e step time with / without
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nrc.py tests/test_gpu_loss_curve.py -x -q > gpurun_out/pytest_host.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_host.log
for d in 1 0 1 0; do NRCHPM_E2E_FRACS=uniform NRCHPM_E2E_DEFER_EMA=$d timeout 120 python scripts/e2e_probe.py 2>&1 | tail -1 | sed "s/^/defer_ema=$d uniform /"; done | tee gpurun_out/e2e_defer.jsonl
NRCHPM_E2E_DEFER_EMA=1 timeout 120 python scripts/e2e_probe.py 2>&1 | tail -1 | sed "s/^/defer_ema=1 graded /" | tee -a gpurun_out/e2e_defer.jsonl
NRCHPM_E2E_TRACE=1 NRCHPM_E2E_FRACS=uniform timeout 120 python scripts/e2e_probe.py 2>gpurun_out/e2e_trace.err | tail -1; grep e2e_trace_us gpurun_out/e2e_trace.err | sed -n '30p'
echo done
