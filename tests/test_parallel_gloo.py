"""CPU, world_size 2 over gloo: the N > 1 path's host logic -- gradient averaging across ranks reproduces the full-batch
gradient (data-parallel training), replicas that apply the same averaged gradient stay bit-identical, tiles partition the screen."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import oracle as O
    from nrc_hpm_renderer_b200.parallel import average_gradients, column_strips
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(11)
    B = 256
    rec = rng.random((B, 5), dtype=np.float32); tgt = (rng.random((B, 3), dtype=np.float32) * 2).astype(np.float32)
    cfg = O.nrc_config(2, 0, 3)                      # TriangleWave + OneBlob: no encoding parameters, deterministic gradients
    m = O.NrcOracle(cfg)
    half = B // world
    m.training_step(rec[rank * half:(rank + 1) * half], tgt[rank * half:(rank + 1) * half], run_optimizer=False)
    g = torch.from_numpy(m.get(m.GRAD).copy())
    average_gradients([g], world)
    full = O.NrcOracle(cfg)
    full.training_step(rec, tgt, run_optimizer=False)
    gf = full.get(full.GRAD)
    # the loss normalises by the batch size (n_total = B*3), so the mean of the two half-batch gradients equals the full-batch one
    err = np.abs(g.numpy() - gf).max() / np.abs(gf).max()
    gathered = [torch.zeros_like(g) for _ in range(world)]
    dist.all_gather(gathered, g)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    strips = column_strips(1920, world)
    q.put((rank, float(err), same, strips))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_gradient_average_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, same, strips in res:
        assert err <= 2e-3, err                      # fp16 rounding of the two half-batch gradients
        assert same                                  # every replica holds the identical averaged gradient
        assert strips == [(0, 960), (960, 1920)]
