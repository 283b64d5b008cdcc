#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python scripts/train_profile.py 2 > gpurun_out/train_profile.jsonl 2> gpurun_out/train_profile.err; echo "profile rc=$?"; cat gpurun_out/train_profile.jsonl; tail -3 gpurun_out/train_profile.err
echo done
