"""GPU, world >= 2 (torchrun): the sharded data-parallel optimizer step over peer memory against NCCL, and replica consistency.
Prints JSON lines on rank 0; exit code 1 on mismatch.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/check_peer_exchange.py

(1) step-by-step form (NRCHPM_PEER_FUSED=0: nrc_peer_gather_kernel -> nrc_adam_kernel on the slice -> nrc_peer_publish_kernel): after
    nrc_peer_exchange every rank holds, on ITS slice of the hash grid, the sum of all ranks' local gradients -- bit for bit the NCCL
    sum at two ranks, to fp16 rounding beyond (the summation order differs) -- and zeros everywhere else; after nrc_optimizer_step
    the fp16 working weights are identical on all ranks.
(2) the default fused form (nrc_peer_adam_kernel: all three in one kernel) gives the SAME bits as the step-by-step form: both train
    on records whose hash-grid gradient is order-independent (copies of a few far-apart records, so the fp16 atomics add equal
    values; 12 hash-grid levels, because on levels 12-15 tcnn's wrapped index arithmetic maps several corners of ONE record onto the
    same entry, whose different weights then add in any order), and working / EMA weights, the owner's master weights and the loss
    are compared bit for bit.
(3) replicas that train on random records for 8 steps hold bit-identical working and EMA weights, and their gradient buffers are clean."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from bench import synth_records, SKY_HALF
from nrc_hpm_renderer_b200 import AppConfig
from nrc_hpm_renderer_b200 import nrc as N
from nrc_hpm_renderer_b200.parallel import PeerGradientExchange, peer_slice

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B = 4096
ok = True


def make(fused, n_levels=16):
    os.environ["NRCHPM_PEER_FUSED"] = "1" if fused else "0"           # read when the cache is created
    app = AppConfig.default()
    cfg = app.model_json()
    cfg["encoding"]["nested"][0]["n_levels"] = n_levels
    cfg.update(infer_batch_size=app.infer_batch_size, train_batch_size=app.train_batch_size, train_batch_count=app.train_batch_count)
    c = N.NeuralRadianceCache(config_json=cfg)
    return c, PeerGradientExchange(c, world)


def all_equal(t):
    gl = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gl, t)
    return all(torch.equal(gl[0], x) for x in gl)


def agree(flag):
    f = torch.tensor([int(flag)], device="cuda"); dist.all_reduce(f, op=dist.ReduceOp.MIN)
    return bool(f.item())


# ---- (1) step-by-step form against NCCL
c, peer = make(False)
n_mlp, n_enc = c.n_mlp_params, c.n_params - c.n_mlp_params
sb, se = peer_slice(n_enc, rank, world)
rng = np.random.default_rng(100 + rank)                      # every rank trains on its own records
for step in range(3):
    rec = torch.from_numpy(synth_records(rng, B)).cuda(); tgt = torch.from_numpy((rng.random((B, 3), dtype=np.float32) * 2).astype(np.float32)).cuda()
    c.training_step(rec, tgt, B, False)
    torch.cuda.synchronize(); dist.barrier()
    local_g = torch.from_numpy(c.get_params(N.GRAD)).cuda()[n_mlp:].to(torch.float16)
    ref = local_g.clone(); dist.all_reduce(ref, op=dist.ReduceOp.SUM)
    peer.run()
    torch.cuda.synchronize(); dist.barrier()
    got = torch.from_numpy(c.get_params(N.GRAD)).cuda()[n_mlp:].to(torch.float16)
    same = agree(torch.equal(got[sb:se], ref[sb:se]))
    maxd = float((got[sb:se].float() - ref[sb:se].float()).abs().max())
    close = agree(maxd <= 2e-3 * float(ref.float().abs().max()))
    cleared = agree(bool((got[:sb] == 0).all() and (got[se:] == 0).all()))
    c.optimizer_step()
    torch.cuda.synchronize(); dist.barrier()
    identical = all_equal(torch.from_numpy(c.get_params(N.WORKING)).cuda())
    if rank == 0:
        print(json.dumps({"step": step, "own_slice_equals_nccl_sum_bitwise": same, "max_abs_diff_rank0": maxd, "consumed_peer_words_cleared": cleared,
                          "working_weights_identical_on_all_ranks": identical, "nonzero_entries_in_sum": int((ref != 0).sum())}))
    ok = ok and cleared and identical and close
c.Destroy()

# ---- (2) fused kernel == step-by-step form, bit for bit, on order-independent gradients
def det_records(r, n_distinct=4, copies=1024):
    base = np.zeros((n_distinct, 5), np.float32)
    cells = r.permutation(8)[:n_distinct]                     # distinct octants of the unit cube: no shared cell corner on any dense level
    base[:, 0] = (cells % 2) / 2 + 0.1 + 0.2 * r.random(n_distinct); base[:, 1] = ((cells // 2) % 2) / 2 + 0.1 + 0.2 * r.random(n_distinct)
    base[:, 2] = (cells // 4) / 2 + 0.1 + 0.2 * r.random(n_distinct)
    base[:, 3:] = r.random((n_distinct, 2))
    return np.repeat(base, copies, axis=0).astype(np.float32)


results = []
modes = [bool(int(v)) for v in os.environ.get("CHECK_MODES", "0,1").split(",")]      # debugging aid: e.g. 1,1 = the fused form twice
for fused in modes:
    c2, peer2 = make(fused, n_levels=12)
    n_mlp2 = c2.n_mlp_params; sb2, se2 = peer_slice(c2.n_params - n_mlp2, rank, world)
    r2 = np.random.default_rng(7 + rank)
    losses = []
    for step in range(4):
        rec = det_records(r2); tgt = np.repeat((r2.random((4, 3)) * 2).astype(np.float32), 1024, axis=0)
        c2.training_step(torch.from_numpy(rec).cuda(), torch.from_numpy(tgt).cuda(), len(rec), False); peer2.run(); c2.optimizer_step()
        losses.append(c2.GetLoss())
    torch.cuda.synchronize(); dist.barrier()
    g = c2.get_params(N.GRAD)[n_mlp2:]
    results.append((c2.get_params(N.WORKING), c2.get_params(N.EMA), c2.get_params(N.MASTER)[n_mlp2 + sb2:n_mlp2 + se2], losses, bool((g == 0).all())))
    c2.Destroy()
a, b = results
fused_equal = agree(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and a[3] == b[3])
clean = agree(a[4] and b[4])
if rank == 0:
    d = {"working": int((a[0] != b[0]).sum()), "ema": int((a[1] != b[1]).sum()), "master_own_slice": int((a[2] != b[2]).sum()), "working_mlp": int((a[0][:n_mlp2] != b[0][:n_mlp2]).sum()),
         "max_abs_working": float(np.abs(a[0] - b[0]).max()), "losses_a": a[3]}
    print(json.dumps({"differing_elements": d}))
    print(json.dumps({"fused_kernel_equals_step_by_step_bitwise": fused_equal, "gradient_buffers_clean_after_step": clean, "losses": b[3]}))
ok = ok and fused_equal and clean

# ---- (3) replicas stay bit-identical (default fused form, random records)
c3, peer3 = make(True)
for step in range(8):
    rec = torch.from_numpy(synth_records(rng, B)).cuda(); tgt = torch.from_numpy((rng.random((B, 3), dtype=np.float32) * 2).astype(np.float32)).cuda()
    c3.training_step(rec, tgt, B, False); peer3.run(); c3.optimizer_step()
torch.cuda.synchronize()
for which, name in ((N.WORKING, "working"), (N.EMA, "ema")):
    same = all_equal(torch.from_numpy(c3.get_params(which)).cuda())
    ok = ok and same
    if rank == 0:
        print(json.dumps({"after_8_steps": name, "replicas_bit_identical": same, "loss": c3.GetLoss()}))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
