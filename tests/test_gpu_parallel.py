"""GPU, needs >= 2 devices on one node (skipped otherwise): the sharded data-parallel optimizer step through peer memory
(nrc_peer_exchange -> nrc_optimizer_step: reduce-scatter, Adam on the rank's slice, weight all-gather) -- every rank's slice of the
summed gradient equals an NCCL all-reduce of the same gradients bit for bit, consumed peer words are cleared, and replicas stay
bit-identical (scripts/check_peer_exchange.py under torchrun)."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def test_peer_exchange_matches_nccl_world2():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one node")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "scripts", "check_peer_exchange.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert '"own_slice_equals_nccl_sum_bitwise": true' in res.stdout and '"replicas_bit_identical": true' in res.stdout
    assert '"fused_kernel_equals_step_by_step_bitwise": true' in res.stdout
