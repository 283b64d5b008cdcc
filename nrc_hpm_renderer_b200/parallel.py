"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch; gloo in the CPU tests).

The hot path shards by screen tiles: tracking and cache inference need no exchange at all; the only exchange
step is the gradient of each training step, averaged over the ranks before Adam + EMA run redundantly (and identically) on
every replica.  The MLP gradient is 24 576 fp32 values; with the hash grid the encoding gradient adds 14.2 M fp16 values.

Only the SETUP lives here (the cudaIpc handles travel once through torch.distributed).  The per-frame schedule -- the frame's
inference cut into chunks that run underneath the gradient exchanges -- is C++ inside the library: after ``PeerGradientExchange``
``nrc.InferAndTrain`` / ``nrc_infer_and_train`` picks it by itself (csrc/nrc.cu: NrcCache::infer_and_train_overlapped).
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
        if b >= e:
            raise ValueError(f"{width} columns cannot be cut into {world} non-empty strips of multiples of {align} columns; use a smaller `align`")
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
            if enc_ptr and n_enc:                 # the large exchange goes first: whatever overlaps the step overlaps this one
                ts.insert(0, torch.as_tensor(_DeviceArray(enc_ptr, n_enc, "<f2"), device="cuda"))
            self._tensors = ts
            self.bytes_per_step = sum(t.numel() * t.element_size() for t in ts)
        average_gradients(self._tensors, self.world, self.group)


def peer_slice(n_encoding_params: int, rank: int, world: int):
    """[begin, end) of the hash-grid PARAMETERS rank `rank` owns in the sharded optimizer step (NrcCache::peer_slice: 16-byte words,
    rounded to whole optimizer warps)"""
    n_vec = n_encoding_params // 8
    per = -(-n_vec // world)
    per = -(-per // 32) * 32
    b = min(n_vec, per * rank)
    return 8 * b, 8 * min(n_vec, b + per)


class PeerGradientExchange:
    """Data-parallel training without a library collective (the cudaIpc handles travel once, at construction, through torch.distributed):
    ``run()`` = nrc_peer_exchange, a reduce-scatter kernel of libnrchpm_b200 over NVLink peer memory that leaves the gradient SUM of
    this rank's slice of the hash grid in its own buffer; the following ``nrc.optimizer_step()`` updates that slice (and the network
    weights, identically on every rank) and all-gathers the new fp16 weights, so replicas hold bit-identical working / EMA weights."""

    def __init__(self, nrc, world: int, group=None):
        import torch.distributed as dist
        self.nrc, self.world = nrc, world
        mine = nrc.peer_export()
        gathered = [None] * world
        dist.all_gather_object(gathered, mine, group=group)
        nrc.peer_setup(dist.get_rank(group), world, b"".join(gathered))
        n_mlp = nrc.n_mlp_params
        n_enc = nrc.n_params - n_mlp
        # bytes a rank moves over NVLink per step: its gradient slice read from every peer, its weight slice written to every peer, plus
        # every peer's MLP gradient (zeros written back over consumed gradient words come on top: <= the touched share)
        self.bytes_per_step = 2 * (world - 1) * (2 * n_enc // world) + (world - 1) * 4 * n_mlp

    def run(self, stream=None):
        self.nrc.peer_exchange(stream)


def make_gradient_exchange(nrc, world: int, group=None, kind: str | None = None):
    """`peer` (default): our own kernel over peer memory; `nccl`: torch.distributed all-reduce"""
    import os
    kind = kind or os.environ.get("NRCHPM_EXCHANGE", "peer")
    if kind == "peer":
        # every rank must take the same path: agree on whether the cudaIpc setup worked everywhere, else use NCCL
        import torch
        import torch.distributed as dist
        ex, err = None, None
        try:
            ex = PeerGradientExchange(nrc, world, group)
        except Exception as e:          # e.g. no peer access between the devices
            err = e
        ok = torch.tensor([1 if ex is not None else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 1:
            return ex
        if dist.get_rank(group) == 0:
            print(f"[nrc_hpm_renderer_b200] peer-memory gradient exchange unavailable ({err}); using the NCCL all-reduce", file=__import__("sys").stderr)
    return GradientAllReduce(nrc, world, group)
