#!/bin/bash
# 1-GPU: timeline of nrc_infer_and_train_host (NRCHPM_E2E_TRACE), graded and uniform chunks
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for g in 1 0; do NRCHPM_E2E_TRACE=1 NRCHPM_E2E_GRADED=$g timeout 120 python scripts/e2e_probe.py 2>gpurun_out/e2e_trace_g$g.err | tail -1; grep e2e_trace_us gpurun_out/e2e_trace_g$g.err | sed -n '30p;31p'; done
echo done
