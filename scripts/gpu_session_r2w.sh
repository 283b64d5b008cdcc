#!/bin/bash
# 1-GPU: path-regeneration tracker -- bit-exactness against the pixel-per-thread kernel, then the sweep of rounds / spill threshold / occupancy
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_wavefront.py -x -q > gpurun_out/pytest_wavefront.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_wavefront.log
for c in 2 4; do
  timeout 120 python scripts/tune_wavefront.py 1 $c 2>&1 | tail -1
  for combo in "1 0 64" "2 16 64" "3 8 64" "3 16 64" "3 24 64" "4 16 64" "4 24 64" "3 16 4" "3 16 6"; do
    set -- $combo
    NRCHPM_WF_ROUNDS=$1 NRCHPM_WF_SPILL_BELOW=$2 NRCHPM_WF_BLOCKS_PER_SM=$3 timeout 120 python scripts/tune_wavefront.py 2 $c 2>&1 | tail -1
  done
done | tee gpurun_out/tune_wavefront.jsonl
M=gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed.sum
timeout 300 ncu --metrics $M --clock-control none -s 12 -c 4 --csv --log-file gpurun_out/launches_wavefront_c2.csv python scripts/tune_wavefront.py 2 2 1 > /dev/null 2>&1
timeout 300 ncu --metrics $M --clock-control none -s 12 -c 4 --csv --log-file gpurun_out/launches_wavefront_c4.csv python scripts/tune_wavefront.py 2 4 1 > /dev/null 2>&1
timeout 300 ncu --metrics $M --clock-control none -s 3 -c 1 --csv --log-file gpurun_out/launches_mono_c4.csv python scripts/tune_wavefront.py 1 4 1 > /dev/null 2>&1
echo done
