#!/bin/bash
# 1-GPU: graded chunk sizes in nrc_infer_and_train_host -- parity of the host entry points, e2e step time with and without
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nrc.py -x -q -k "host or edge or ragged" > gpurun_out/pytest_host.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_host.log
for g in 1 0 1 0; do NRCHPM_E2E_GRADED=$g timeout 120 python scripts/e2e_probe.py 2>&1 | tail -1 | sed "s/^/graded=$g /"; done | tee gpurun_out/e2e_graded.jsonl
echo done
