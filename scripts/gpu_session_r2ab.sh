#!/bin/bash
# 1-GPU: gen_rays pass under register / occupancy variants of the lookup diet (and the pre-diet kernel)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in oldlookup_nolut_mb9 oldlookup_nolut_mb12 oldlookup_mb9 oldlookup_mb12 intfloor_nolut_mb9 intfloor_nolut_mb12; do
  for c in 2 4; do NRCHPM_LIB=$PWD/nrc_hpm_renderer_b200/variants/libnrchpm_b200_$v.so timeout 120 python scripts/tune_wavefront.py 1 $c 2>&1 | tail -1 | sed "s/^/$v /"; done
done | tee gpurun_out/tune_tracker_variants.jsonl
echo done
