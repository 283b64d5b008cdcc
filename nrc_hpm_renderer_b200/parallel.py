"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch; gloo in the CPU tests).

The hot path shards by screen tiles: tracking and cache inference need no exchange at all; the only exchange
step is the gradient of each training step, averaged over the ranks before Adam + EMA run redundantly (and identically) on
every replica.  The MLP gradient is 24 576 fp32 values; with the hash grid the encoding gradient adds 14.2 M fp16 values.
"""
from __future__ import annotations

import numpy as np


class _DeviceArray:
    """minimal __cuda_array_interface__ carrier so torch can wrap a raw device pointer owned by the C library"""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}


def column_strips(width: int, world: int, align: int = 64):
    """Column strips [x_begin, x_end) per rank.  Records are indexed x*H + y, so a strip is one contiguous slice of the
    inference buffers; strips are multiples of `align` columns except the last."""
    per = -(-width // world)
    per = -(-per // align) * align
    out = []
    for r in range(world):
        b, e = min(width, r * per), min(width, (r + 1) * per)
        out.append((b, e))
    return out


def average_gradients(tensors, world: int, group=None):
    """all-reduce (mean) a list of torch tensors in place; works on NCCL (GPU) and gloo (CPU tests)"""
    import torch.distributed as dist
    for t in tensors:
        if dist.get_backend(group) == "nccl":
            dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            t.div_(world)
    return tensors


class GradientAllReduce:
    """Averages the gradients of the last nrc.training_step(run_optimizer=False) across ranks."""

    def __init__(self, nrc, world: int, group=None):
        self.nrc, self.world, self.group = nrc, world, group
        self._tensors = None
        self.bytes_per_step = 0

    def run(self):
        import torch
        mlp_ptr, enc_ptr = self.nrc.gradient_buffers()       # collapses the per-chunk partials into one fp32 buffer
        if self._tensors is None:
            n_mlp = self.nrc.n_mlp_params
            ts = [torch.as_tensor(_DeviceArray(mlp_ptr, n_mlp, "<f4"), device="cuda")]
            n_enc = self.nrc.n_params - n_mlp
            if enc_ptr and n_enc:
                ts.append(torch.as_tensor(_DeviceArray(enc_ptr, n_enc, "<f2"), device="cuda"))
            self._tensors = ts
            self.bytes_per_step = sum(t.numel() * t.element_size() for t in ts)
        average_gradients(self._tensors, self.world, self.group)
