#!/bin/bash
# 1-GPU: e2e step time under different chunk boundaries of nrc_infer_and_train_host (after moving the loss read-back off the critical path)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nrc.py -x -q -k "host" > gpurun_out/pytest_host.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_host.log
for f in uniform "0.40,0.65,0.82,0.93" "0.6,0.87" "0.55,0.8,0.93" "0.65" "0.5,0.75,0.9" "0.45,0.8" uniform "0.6,0.87"; do NRCHPM_E2E_FRACS=$f timeout 120 python scripts/e2e_probe.py 2>&1 | tail -1 | sed "s/^/fracs=$f /"; done | tee gpurun_out/e2e_fracs.jsonl
NRCHPM_E2E_TRACE=1 NRCHPM_E2E_FRACS="0.6,0.87" timeout 120 python scripts/e2e_probe.py 2>gpurun_out/e2e_trace.err | tail -1; grep e2e_trace_us gpurun_out/e2e_trace.err | sed -n '30p'
echo done
