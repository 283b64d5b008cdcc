#!/bin/bash
# 1-GPU: e2e (host buffers through nrc_infer_and_train_host) under the chunk-size and concurrent-inference knobs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for t in 8 4 2; do
  for s in 0 88 100 112 124; do
    NRCHPM_E2E_CHUNK_TILES=$t NRCHPM_E2E_CONCURRENT_SMS=$s timeout 120 python scripts/e2e_probe.py 2>&1 | tail -1 | sed "s/^/sms=$s /"
  done
done | tee gpurun_out/e2e_knobs.jsonl
echo done
