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
            if enc_ptr and n_enc:                 # the large exchange goes first: whatever overlaps the step overlaps this one
                ts.insert(0, torch.as_tensor(_DeviceArray(enc_ptr, n_enc, "<f2"), device="cuda"))
            self._tensors = ts
            self.bytes_per_step = sum(t.numel() * t.element_size() for t in ts)
        average_gradients(self._tensors, self.world, self.group)


class PeerGradientExchange:
    """Same contract as GradientAllReduce.run(), without a library collective: one kernel of libnrchpm_b200 sums the gradients
    across the ranks through peer memory over NVLink (nrc_peer_exchange; the cudaIpc handles travel once, at construction, through
    torch.distributed).  Afterwards nrc.optimizer_step() applies the mean gradient on every replica."""

    def __init__(self, nrc, world: int, group=None):
        import torch.distributed as dist
        self.nrc, self.world = nrc, world
        mine = nrc.peer_export()
        gathered = [None] * world
        dist.all_gather_object(gathered, mine, group=group)
        nrc.peer_setup(dist.get_rank(group), world, b"".join(gathered))
        n_mlp = nrc.n_mlp_params
        n_enc = nrc.n_params - n_mlp
        # bytes a rank moves over NVLink per step, upper bound (untouched entries are not stored): own slice read from and
        # written to every peer, plus every peer's MLP gradient
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


class OverlappedInferAndTrain:
    """``InferAndTrain`` for one screen tile per rank with the gradient exchange hidden ("overlap with inference of
    the tile ... and report exposed time").

    The reference runs Inference() and then Train() (src/NeuralRadianceCache.cu:97-156); with data-parallel training every
    training step waits for an all-reduce of the gradients (96 KB fp32 MLP + 28.5 MB fp16 hash grid) during which the SMs idle.
    Here the tile's inference is cut into one chunk per training step and each chunk is released on a second, lower-priority
    stream at the moment the step's backward pass has finished -- it runs underneath the all-reduce (NCCL needs a few SMs) and the
    memory-bound Adam + EMA pass.  The cache is evaluated from a snapshot of the pre-training parameters (nrc_snapshot_params), so
    the result is the reference's: Inference() sees the weights of the previous frame."""

    def __init__(self, nrc, world: int, group=None, n_records_align: int = 128):
        import torch
        self.nrc, self.world = nrc, world
        self.allreduce = make_gradient_exchange(nrc, world, group) if world > 1 else None
        self.s_inf = torch.cuda.Stream(priority=0)
        self.s_train = torch.cuda.Stream(priority=-1)
        self.align = n_records_align
        # persistent-grid cap of the chunks that ride along with an all-reduce: one CTA per SM leaves registers and shared memory
        # for NCCL's CTAs on every SM (with two per SM the all-reduce kernel waits until the chunk has drained)
        import os
        sm = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
        self.cta_limit = int(os.environ.get("NRCHPM_OVERLAP_CTAS", sm)) if world > 1 else int(os.environ.get("NRCHPM_OVERLAP_CTAS", 0))
        # world == 1: nothing to hide behind; "whole" releases the tile's inference as ONE capped launch next to the training steps
        self.whole = world == 1 and os.environ.get("NRCHPM_OVERLAP_WHOLE", "0") == "1"
        # share of the tile's records that rides along with the exchanges (capped, slower launches); the rest is evaluated up
        # front at full occupancy.  Sized so that a capped chunk takes about as long as one exchange + optimizer window.
        self.window_fraction = float(os.environ.get("NRCHPM_OVERLAP_FRACTION", "0.6" if world > 1 else "1.0"))
        self._events = [torch.cuda.Event() for _ in range(64)]
        self._ev_i = 0

    def _event(self):
        e = self._events[self._ev_i % len(self._events)]
        self._ev_i += 1
        return e

    def run(self, d_in, d_out, n: int, d_train_in, d_train_target, batch: int, n_batches: int):
        import torch
        from .nrc import SNAPSHOT
        nrc, s_inf, s_tr = self.nrc, self.s_inf, self.s_train
        cur = torch.cuda.current_stream()
        s_inf.wait_stream(cur); s_tr.wait_stream(cur)
        nrc.snapshot_params(True, s_inf.cuda_stream)
        e = self._event(); e.record(s_inf); s_tr.wait_event(e)            # training overwrites what the snapshot copy reads
        off = 0
        if self.window_fraction < 1.0 and not self.whole and n_batches > 0:
            head = int(n * (1.0 - self.window_fraction)) // self.align * self.align
            if head > 0:                                                    # full-occupancy launch first; training starts behind it
                nrc.inference(d_in[:head], d_out[:head], head, SNAPSHOT, s_inf.cuda_stream)
                e = self._event(); e.record(s_inf); s_tr.wait_event(e)
                off = head
        chunk = -(-(n - off) // max(n_batches, 1))
        chunk = -(-chunk // self.align) * self.align
        if self.whole:
            nrc.set_inference_cta_limit(self.cta_limit)
            nrc.inference(d_in[:n], d_out[:n], n, SNAPSHOT, s_inf.cuda_stream)
            nrc.set_inference_cta_limit(0)
            off = n
        with torch.cuda.stream(s_tr):                                       # NCCL orders itself behind the current stream
            for b in range(n_batches):
                nrc.training_step(d_train_in[b * batch:(b + 1) * batch], d_train_target[b * batch:(b + 1) * batch], batch, False, s_tr.cuda_stream)
                e = self._event(); e.record(s_tr)                          # backward + weight gradients of step b are done
                if self.allreduce is not None:                              # queued FIRST: its CTAs must not wait for the chunk's
                    self.allreduce.run(s_tr.cuda_stream) if isinstance(self.allreduce, PeerGradientExchange) else self.allreduce.run()
                s_inf.wait_event(e)                                         # chunk b rides along with all-reduce b and optimizer b
                m = min(chunk, n - off) if b < n_batches - 1 else n - off
                if m > 0:
                    nrc.set_inference_cta_limit(self.cta_limit)
                    nrc.inference(d_in[off:off + m], d_out[off:off + m], m, SNAPSHOT, s_inf.cuda_stream)
                    nrc.set_inference_cta_limit(0)
                    off += m
                nrc.optimizer_step(s_tr.cuda_stream)
        if off < n:                                                         # no training this frame: plain Inference()
            nrc.inference(d_in[off:n], d_out[off:n], n - off, SNAPSHOT, s_inf.cuda_stream)
        cur.wait_stream(s_inf); cur.wait_stream(s_tr)
