#!/bin/bash
cd "$(dirname "$0")/.."
echo "== serial schedule timeline (names shifted: inference first)"
NRCHPM_OVERLAP=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/overlap_timeline.py 2>&1 | grep -v "^\*\|OMP_NUM\|NCCL\|^$" | head -14
