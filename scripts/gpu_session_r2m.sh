#!/bin/bash
cd "$(dirname "$0")/.."
for c in 50 60; do
echo "== carveout $c"
NRCHPM_WS_CARVEOUT=$c NRCHPM_OVERLAP_HEAD=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/overlap_timeline.py 2>&1 | grep -v "^\*\|OMP_NUM\|NCCL\|^$" | head -12
done
