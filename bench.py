#!/usr/bin/env python
"""bench.py -- headline benchmark of the NRC-HPM hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): one 1920x1080 frame of the reference's default configuration
(`RelativeL2Luminance Adam 0.01 0.99 0 0 64 6 21 14 4 ...`, reference src/main.cu:432-439): HashGrid16x2 + OneBlob4
encoding, 64 x 6 fully fused MLP.  One "step" is NeuralRadianceCache::InferAndTrain for that frame (reference
src/NeuralRadianceCache.cu:97-156): inference on W*H = 2 073 600 query records + 4 training steps of 2^14 records.
`value` = queries (inference + training records) per second, whole job, inputs resident in HBM.
`e2e`   = the same step through the host-buffer C-ABI entry points (pinned host memory, H2D + D2H inside the timed region).
`frame` = the full frame loop on the bundled cloud (tracking + NRC + compositing) with per-stage milliseconds.
N > 1 (torchrun): weak scaling -- every rank owns one 1080p screen tile (its own query records and training records),
weights are replicated and the gradients of every training step are averaged with an NCCL all-reduce over NVLink.

--impl reference times the reference's own implementation of the step: tiny-cuda-nn built unmodified for sm_100a
(oracle/_ref/tcnn_oracle, driven exactly like en::NeuralRadianceCache); if that binary is absent, the scalar CPU oracle.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 1920, 1080
N_INFER = W * H
TRAIN_BATCH, TRAIN_BATCHES = 1 << 14, 4
N_SETS = 4                                   # record sets rotated between steps: 4 x (41.5 + 24.9) MB > 126 MB L2
SKY_HALF = np.array([62.317, 42.295, 76.707], np.float32) / 2
METRIC, UNIT = "nrc_queries_per_s_1080p_infer_and_train", "queries/s"
FLOP_PER_QUERY_H6 = 2 * (48 * 64 + 5 * 64 * 64 + 64 * 3)        # 47 488 (H = 6)
BYTES_PER_QUERY = 20 + 12 + 16 * 8 * 4                            # record in + radiance out + hash-grid gathers = 544 B
NCU_DRAM_BYTES_PER_LAUNCH = 70_055_168 + 11_916_032               # ncu capture of the inference launch (profiles/, round 1)


def synth_records(rng, n):
    """SURVEY.md 8(d) synthetic inputs: pos ~ U[0,1)^3 + skySize/2 (reference normalisation, Q4), theta ~ U[-.5,1.5), phi ~ U[0,1)"""
    rec = rng.random((n, 5), dtype=np.float32)
    rec[:, :3] += SKY_HALF
    rec[:, 3] = rec[:, 3] * 2 - 0.5
    return rec


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)"""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return {"hbm_gbs": j["hbm_gbs"], "tflops": j["bf16_tflops"], "tflops_sustained": j.get("bf16_tflops_sustained", j["bf16_tflops"]), "which": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1590.0, "tflops_sustained": 1400.0, "which": "fallback"}


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    binp = os.path.join(ROOT, "oracle", "_ref", "tcnn_oracle")
    cfg = {"workload": "NeuralRadianceCache::InferAndTrain, one 1920x1080 frame: 2073600 inference records + 4x16384 training records, "
                       "HashGrid16x2+OneBlob4, MLP 64x6 (reference default argv)", "records": "synthetic, seed 1337", "l2": f"{N_SETS} record sets rotated (> L2)"}
    if os.path.exists(binp):
        cmd = [binp, "bench", f"n_infer={N_INFER}", f"batch={TRAIN_BATCH}", f"batches={TRAIN_BATCHES}", f"frames={args.steps}", f"warmup={max(args.warmup, 3)}",
               "pos=0", "dir=0", "depth=6", f"sets={N_SETS}"]
        sampler = ClockSampler(); sampler.start()
        res = subprocess.run(cmd, capture_output=True, text=True)
        clocks = sampler.stop()
        line = [l for l in res.stdout.splitlines() if l.startswith("{")]
        if res.returncode != 0 or not line:
            print(json.dumps({"impl": "reference", "unavailable": "tcnn_oracle failed: " + (res.stderr.strip().splitlines() or ["?"])[-1][:200]}))
            return 0
        j = json.loads(line[-1])
        v = j["queries_per_s"]
        out = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": j["ms_per_frame"],
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic", "config": cfg, "impl": "reference",
               "cpu_baseline": {"value": v, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "reference's tiny-cuda-nn (unmodified, built for sm_100a) on the same B200; the reference has no CPU implementation of this path"},
               "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "clocks": clocks, "loss": j.get("loss")}
        print(json.dumps(out))
        return 0
    # no tcnn build on this box: time the scalar CPU restatement on a bounded sample of the same workload
    import oracle as O
    O.build()
    o = O.NrcOracle(O.nrc_config(0, 0, 6))
    rng = np.random.default_rng(1337)
    n_i, n_t = 16384, 1024
    rec, tin, tgt = synth_records(rng, n_i), synth_records(rng, n_t), (rng.random((n_t, 3), dtype=np.float32) * 2).astype(np.float32)
    times = []
    for s in range(max(args.warmup, 1) + args.steps):
        t0 = time.perf_counter()
        o.inference(rec); o.training_step(tin, tgt)
        times.append(time.perf_counter() - t0)
    t = float(np.mean(times[max(args.warmup, 1):]))
    v = (n_i + n_t) / t
    cores = os.cpu_count() or 1
    print(json.dumps({"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic", "config": cfg, "impl": "reference",
                      "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": f"{n_i} inference + {n_t} training records per step"},
                      "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
    return 0


# ------------------------------------------------------------------------------------------------ our arm
def cpu_baseline_leg():
    import oracle as O
    O.build()
    o = O.NrcOracle(O.nrc_config(0, 0, 6))
    rng = np.random.default_rng(1337)
    n_i, n_t = N_INFER, TRAIN_BATCH            # one full frame of inference records + one 2^14 training step (~10-20 s of host time)
    rec, tin, tgt = synth_records(rng, n_i), synth_records(rng, n_t), (rng.random((n_t, 3), dtype=np.float32) * 2).astype(np.float32)
    t0 = time.perf_counter()
    o.inference(rec)
    o.training_step(tin, tgt)
    t = time.perf_counter() - t0
    cores = int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))
    return {"value": (n_i + n_t) / t, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle/nrc_oracle.cpp (OpenMP, {cores} threads): {n_i} inference records + one {n_t}-record training step, {t:.1f} s"}


def frame_leg(torch, stream, steps, warmup):
    """full frame loop on the bundled cloud: tracking + NRC + compositing (per-stage ms, density-lookup roofline)"""
    from nrc_hpm_renderer_b200 import AppConfig, Camera, HpmSceneConfig, volume
    from nrc_hpm_renderer_b200.nrc import NeuralRadianceCache
    from nrc_hpm_renderer_b200.renderer import BUF_COUNTERS, HpmScene, NrcHpmRenderer
    path = os.path.join(ROOT, "data", "wdas_cloud_quarter_u8.npz")
    if os.path.exists(path):
        grid, vol_name = volume.load_volume(path).data, "wdas_cloud_quarter (498x338x613 u8, bundled cloud)"
    else:
        small = np.load(os.path.join(ROOT, "tests", "golden", "wdas_cloud_sixteenth_u8.npz"))["data"]
        grid, vol_name = np.ascontiguousarray(small.repeat(4, 0).repeat(4, 1).repeat(4, 2)), "wdas_cloud_sixteenth upsampled x4 (quarter fixture missing)"
    app = AppConfig.default()
    app.scene = HpmSceneConfig.preset(0)
    out = {"volume": vol_name, "scene": 0}
    for mode, compact, pipelined in (("compact", True, False), ("all_records", False, False), ("compact_pipelined_training", True, True)):
        nrc = NeuralRadianceCache(app)
        scene = HpmScene(grid, app.scene)
        # pipelined: Train(N) runs underneath the tracking passes of frame N+1 (same order of effects); per-frame time = the main
        # stream from gen_rays to compositing, which includes waiting for the previous frame's training before Inference()
        r = NrcHpmRenderer(W, H, False, Camera(aspect=W / H), app, scene, nrc, compact_inference=compact, pipeline_train=pipelined, stream=stream)
        rng = np.random.default_rng(1337)
        stages = []
        for i in range(warmup + steps):
            r.Render(True, rng.random(4).astype(np.float32))
            if i >= warmup:
                stages.append(r.EvaluateTimestampQueries())
        cnt = r.read(BUF_COUNTERS)
        ms = {k: float(np.mean([s[k] for s in stages])) for k in stages[0]}
        out[mode] = {"ms": {k: round(v, 4) for k, v in ms.items()}, "frames_per_s": 1e3 / ms["total"], "density_lookups_gen_rays": int(cnt[0]),
                     "density_lookups_prep_train": int(cnt[1]), "active_records": int(cnt[2]), "loss": nrc.GetLoss()}
        if compact and not pipelined:
            # tracking roofline: L lookups x 1 B + P pixels x 36 B, and the sector-granular figure L x 32 B
            L, P = int(cnt[0]), W * H
            t = ms["gen_rays"] * 1e-3
            out["tracking_roofline"] = {"algorithmic_GBps": (L + 36 * P) / t / 1e9, "sector_GBps": (32 * L + 36 * P) / t / 1e9, "lookups_per_s": L / t}
        r.Destroy(); scene.Destroy(); nrc.Destroy()
    return out


def frame_tiles_leg(torch, dist, rank, world, steps, warmup):
    """BASELINE config 5: a 3840x2160 frame cut into `world` column strips (one per GPU), tracking + inference per strip without any
    exchange, cache training data-parallel (1/world of every batch per rank, gradients summed through peer memory).  ms per frame =
    max over ranks of the per-rank frame time (CUDA events)."""
    from nrc_hpm_renderer_b200 import AppConfig, Camera, HpmSceneConfig, volume
    from nrc_hpm_renderer_b200.nrc import NeuralRadianceCache
    from nrc_hpm_renderer_b200.parallel import PeerGradientExchange
    from nrc_hpm_renderer_b200.renderer import HpmScene, NrcHpmRenderer, make_tile_render_config, tile_app_config
    Wt, Ht = 3840, 2160
    path = os.path.join(ROOT, "data", "wdas_cloud_quarter_u8.npz")
    if not os.path.exists(path):
        return {"error": "bundled volume fixture missing"}
    grid = volume.load_volume(path).data
    app = AppConfig.default(); app.scene = HpmSceneConfig.preset(0)
    ta = tile_app_config(app, world)
    nrc = NeuralRadianceCache(ta)
    PeerGradientExchange(nrc, world)                       # Train() inside Render() now exchanges after every step
    scene = HpmScene(grid, ta.scene)
    cfg = make_tile_render_config(Wt, Ht, ta, rank, world)
    r = NrcHpmRenderer(Wt, Ht, False, Camera(aspect=Wt / Ht), ta, scene, nrc, render_config=cfg)
    rng = np.random.default_rng(1337)
    ms = []
    for i in range(warmup + steps):
        fr = rng.random(4).astype(np.float32)
        dist.barrier()
        r.Render(True, fr); r.sync()
        if i >= warmup:
            ms.append(r.GetFrameTimeMS())
    t = torch.tensor([float(np.mean(ms))], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out = {"resolution": [Wt, Ht], "tiles": world, "strip_columns": [int(cfg.x_begin), int(cfg.x_end)], "train_records_per_rank_and_step": ta.train_batch_size,
           "ms_per_frame": float(t.item()), "frames_per_s": 1e3 / float(t.item()), "loss": nrc.GetLoss()}
    r.Destroy(); scene.Destroy(); nrc.Destroy()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path is sm_100a code with no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from nrc_hpm_renderer_b200 import AppConfig, _lib
    from nrc_hpm_renderer_b200.nrc import NeuralRadianceCache
    from nrc_hpm_renderer_b200.parallel import GradientAllReduce, OverlappedInferAndTrain, PeerGradientExchange, make_gradient_exchange

    app = AppConfig.default()                          # reference default argv: hash grid + OneBlob, 64 x 6, lr 0.01, EMA 0.99
    nrc = NeuralRadianceCache(app)
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream
    rng = np.random.default_rng(1337 + rank)
    d_in = [torch.from_numpy(synth_records(rng, N_INFER)).cuda() for _ in range(N_SETS)]
    d_out = [torch.empty((N_INFER, 3), dtype=torch.float32, device="cuda") for _ in range(N_SETS)]
    n_train = TRAIN_BATCH * TRAIN_BATCHES
    h_tin = [synth_records(rng, n_train) for _ in range(N_SETS)]
    h_tgt = [(rng.random((n_train, 3), dtype=np.float32) * 2).astype(np.float32) for _ in range(N_SETS)]
    d_tin = [torch.from_numpy(a).cuda() for a in h_tin]
    d_tgt = [torch.from_numpy(a).cuda() for a in h_tgt]
    overlap_default = 1 if world > 1 else 0          # measured at N = 2: 1.53 ms overlapped vs 1.61 ms serial (profiles/r01_summary.md)
    # N > 1: the tile's inference runs underneath the gradient all-reduces (parallel.OverlappedInferAndTrain); N = 1 keeps the
    # reference's serial Inference() -> Train() schedule unless --overlap asks for the same two-stream schedule
    overlap = args.overlap if args.overlap is not None else overlap_default
    runner = OverlappedInferAndTrain(nrc, world) if overlap else None
    allreduce = make_gradient_exchange(nrc, world) if (world > 1 and runner is None) else None

    def exchange_gradients():
        if isinstance(allreduce, PeerGradientExchange):
            allreduce.run(sp)
        else:
            allreduce.run()

    def step(i):
        s = i % N_SETS
        if runner is not None:
            runner.run(d_in[s], d_out[s], N_INFER, d_tin[s], d_tgt[s], TRAIN_BATCH, TRAIN_BATCHES)
            return
        nrc.inference(d_in[s], d_out[s], N_INFER, True, sp)                      # Inference(): one batch (2^21 >= W*H), EMA weights
        for b in range(TRAIN_BATCHES):                                           # Train(): 4 x training_step
            tin = d_tin[s][b * TRAIN_BATCH:(b + 1) * TRAIN_BATCH]; tgt = d_tgt[s][b * TRAIN_BATCH:(b + 1) * TRAIN_BATCH]
            if allreduce is None:
                nrc.training_step(tin, tgt, TRAIN_BATCH, True, sp)
            else:
                nrc.training_step(tin, tgt, TRAIN_BATCH, False, sp)
                exchange_gradients()
                nrc.optimizer_step(sp)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    warmup = max(args.warmup, 3)
    for i in range(warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.lib().nrchpm_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev_k = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0.record(stream)
    for i in range(args.steps):
        s = (warmup + i) % N_SETS
        if runner is not None:
            step(warmup + i)
            continue
        # dominant kernel, timed live on its own stream: the fused encode + MLP inference launch
        ev_k[i][0].record(stream)
        nrc.inference(d_in[s], d_out[s], N_INFER, True, sp)
        ev_k[i][1].record(stream)
        for b in range(TRAIN_BATCHES):
            tin = d_tin[s][b * TRAIN_BATCH:(b + 1) * TRAIN_BATCH]; tgt = d_tgt[s][b * TRAIN_BATCH:(b + 1) * TRAIN_BATCH]
            if allreduce is None:
                nrc.training_step(tin, tgt, TRAIN_BATCH, True, sp)
            else:
                nrc.training_step(tin, tgt, TRAIN_BATCH, False, sp)
                exchange_gradients()
                nrc.optimizer_step(sp)
    e1.record(stream)
    barrier()
    launches = _lib.lib().nrchpm_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    if runner is not None:
        # the overlapped schedule has no serial inference launch to bracket: time the dominant kernel alone afterwards
        for i in range(args.steps):
            s = (warmup + i) % N_SETS
            ev_k[i][0].record(stream)
            nrc.inference(d_in[s], d_out[s], N_INFER, True, sp)
            ev_k[i][1].record(stream)
        torch.cuda.synchronize()
    ms_kernel = float(np.mean([a.elapsed_time(b) for a, b in ev_k]))
    # exposed cost of the gradient exchange: the same schedule with the all-reduce left out (replicas diverge; timing only)
    exchange = None
    if world > 1:
        saved = runner.allreduce if runner is not None else None
        if runner is not None:
            runner.allreduce = None
        barrier()
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_x = max(3, min(args.steps, 50))
        x0.record(stream)
        for i in range(n_x):
            if runner is not None:
                step(i)
            else:
                s = i % N_SETS
                nrc.inference(d_in[s], d_out[s], N_INFER, True, sp)
                for b in range(TRAIN_BATCHES):
                    nrc.training_step(d_tin[s][b * TRAIN_BATCH:(b + 1) * TRAIN_BATCH], d_tgt[s][b * TRAIN_BATCH:(b + 1) * TRAIN_BATCH], TRAIN_BATCH, True, sp)
        x1.record(stream)
        barrier()
        ms_nocomm = x0.elapsed_time(x1) / n_x
        if runner is not None:
            runner.allreduce = saved
        ar = runner.allreduce if runner is not None else allreduce
        exchange = {"kind": "own kernel over peer memory (nrc_peer_reduce_kernel: P2P loads/stores over NVLink)" if isinstance(ar, PeerGradientExchange) else "NCCL all-reduce",
                    "bytes_per_training_step": ar.bytes_per_step, "all_reduces_per_step": TRAIN_BATCHES, "ms_per_step_without_exchange": ms_nocomm,
                    "schedule": "inference chunks overlap the all-reduces" if runner is not None else "serial"}
    if world > 1:
        t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    ms_step = ms / args.steps
    q_step = (N_INFER + n_train) * world
    value = q_step / (ms_step * 1e-3)

    # ---- e2e: host buffers through the C ABI (pinned memory), H2D + D2H every step
    pin_in = [torch.from_numpy(synth_records(rng, N_INFER)).pin_memory() for _ in range(2)]
    pin_out = torch.empty((N_INFER, 3), dtype=torch.float32).pin_memory()
    pin_tin = [torch.from_numpy(h_tin[i]).pin_memory() for i in range(2)]
    pin_tgt = [torch.from_numpy(h_tgt[i]).pin_memory() for i in range(2)]

    np_in, np_out = [t.numpy() for t in pin_in], pin_out.numpy()
    np_tin, np_tgt = [t.numpy() for t in pin_tin], [t.numpy() for t in pin_tgt]

    def e2e_step(i):
        s = i % 2
        # one InferAndTrain call on pinned host buffers: H2D of the records, kernels, D2H of the radiance + the loss
        nrc.infer_and_train_host(np_in[s], np_out, np_tin[s], np_tgt[s], TRAIN_BATCH, True)
    for i in range(3):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, min(args.steps, 50))
    for i in range(e2e_steps):
        e2e_step(i)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    if world > 1:
        t = torch.tensor([e2e_ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); e2e_ms = float(t.item())
    e2e_value = q_step / (e2e_ms / e2e_steps * 1e-3)
    h2d = N_INFER * 20 + n_train * 32
    d2h = N_INFER * 12 + 4 * TRAIN_BATCHES

    frame_tiles = None
    if world > 1 and not args.no_frame:
        try:
            frame_tiles = frame_tiles_leg(torch, dist, rank, world, steps=max(3, min(args.steps, 30)), warmup=5)
        except Exception as e:
            frame_tiles = {"error": str(e)[:300]}
    if rank == 0:
        pk = peaks()
        achieved_gbs = BYTES_PER_QUERY * N_INFER / (ms_kernel * 1e-3) / 1e9
        achieved_tf = FLOP_PER_QUERY_H6 * N_INFER / (ms_kernel * 1e-3) / 1e12
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
               "config": {"workload": "NeuralRadianceCache::InferAndTrain, one 1920x1080 frame per GPU: 2073600 inference records + 4x16384 training records, "
                                      "HashGrid16x2+OneBlob4, MLP 64x6 (reference default argv)", "records": "synthetic, seed 1337", "l2": f"{N_SETS} record sets rotated (265 MB > 126 MB L2)",
                          "parallelism": f"tiles x{world}, data-parallel training" if world > 1 else "single GPU"},
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / e2e_steps},
               "gpu_launches": int(launches), "clocks": clocks,
               "roofline": {"kernel": "nrc_forward_kernel<48,false> (fused hash-grid/OneBlob encode + 7-layer tcgen05 MLP + fp32 output)", "bound": "hbm",
                            "achieved": achieved_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved_gbs / pk["hbm_gbs"], "traffic": NCU_DRAM_BYTES_PER_LAUNCH,
                            "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this launch (profiles/r01_ncu_fwd_infer_one_level_per_round.txt): "
                                              "the 28.5 MB hash tables are L2-resident, DRAM sees the record I/O (66.4 MB) only",
                            "l2": {"sectors_per_query": 57.2, "l1_to_l2_request_unit_busy": 0.66, "lts_throughput": 0.49, "note": "ncu, same capture: the unit this kernel loads most is the L1->L2 request path (one 32-byte sector per request, divergent gathers)"},
                            "peak_source": pk["which"], "ms_per_launch": ms_kernel, "algorithmic_bytes_per_query": BYTES_PER_QUERY,
                            "tensor": {"achieved_tflops": achieved_tf, "peak_tflops": pk["tflops_sustained"], "frac": achieved_tf / pk["tflops_sustained"], "flop_per_query": FLOP_PER_QUERY_H6}},
               "loss": nrc.GetLoss()}
        out["config"]["schedule"] = "inference overlapped with training (snapshot of the pre-training parameters)" if runner is not None else "Inference() then Train(), serial (reference order)"
        if exchange is not None:
            exchange["exposed_ms_per_step"] = ms_step - exchange["ms_per_step_without_exchange"]
            out["gradient_exchange"] = exchange
        if frame_tiles is not None:
            out["frame_4k_tiles"] = frame_tiles
        if world == 1:
            try:
                out["cpu_baseline"] = cpu_baseline_leg()
            except Exception as e:          # the oracle is only the checker / baseline; never part of the product path
                out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"unavailable: {e}"}
            if not args.no_frame:
                try:
                    out["frame"] = frame_leg(torch, sp, steps=max(3, min(args.steps, 30)), warmup=5)
                except Exception as e:
                    out["frame"] = {"error": str(e)[:300]}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--overlap", type=int, default=None, help="1: two-stream schedule (inference chunks under the optimizer / all-reduce); default: only for N > 1")
    ap.add_argument("--no-frame", action="store_true", help="skip the full-frame leg (tracking + NRC + compositing)")
    args = ap.parse_args()
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
