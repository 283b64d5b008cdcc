"""GPU, world >= 2 (torchrun): a 4K frame partitioned into column strips, one per rank (SURVEY.md 8e / BASELINE config 5), with
data-parallel cache training through the peer-memory gradient exchange.
Checks: (1) frame 0 (no cache term yet, Q7) stitched from the strips equals the single-GPU frame bit for bit -- tracking and
compositing shard without any exchange; (2) the replicas' parameters stay bit-identical over the frames; (3) later frames are
finite and carry a cache term.  Prints per-frame milliseconds (max over ranks) for N ranks and, on rank 0, for one GPU alone.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 scripts/check_tiles.py [W H frames]
"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from nrc_hpm_renderer_b200 import AppConfig, Camera, HpmSceneConfig, volume
from nrc_hpm_renderer_b200 import nrc as N
from nrc_hpm_renderer_b200.parallel import PeerGradientExchange, column_strips
from nrc_hpm_renderer_b200.renderer import BUF_OUTPUT, HpmScene, NrcHpmRenderer, make_render_config, make_tile_render_config, tile_app_config

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H, FRAMES = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (3840, 2160, 12)
path = os.path.join(ROOT, "data", "wdas_cloud_quarter_u8.npz")
grid = volume.load_volume(path).data if os.path.exists(path) else np.ascontiguousarray(np.load(os.path.join(ROOT, "tests", "golden", "wdas_cloud_sixteenth_u8.npz"))["data"])
app = AppConfig.default(); app.scene = HpmSceneConfig.preset(0)
cam = Camera(aspect=W / H)
tile_app = tile_app_config(app, world)                          # 1/world of every training batch per rank
tile_cfgs = [make_tile_render_config(W, H, tile_app, r_, world) for r_ in range(world)]
strips = [(int(c.x_begin), int(c.x_end)) for c in tile_cfgs]


def run(peer):
    a = tile_app if peer else app
    cache = N.NeuralRadianceCache(a)
    if peer:
        PeerGradientExchange(cache, world)                      # from here on Train() exchanges after every step
    scene = HpmScene(grid, a.scene)
    cfg = tile_cfgs[rank] if peer else make_render_config(W, H, app)
    r = NrcHpmRenderer(W, H, False, cam, a, scene, cache, render_config=cfg)
    rng = np.random.default_rng(1337)
    imgs, ms = [], []
    for f in range(FRAMES):
        fr = rng.random(4).astype(np.float32)
        if peer:
            dist.barrier()
        r.Render(True, fr); r.sync()
        ms.append(r.GetFrameTimeMS())
        if f in (0, FRAMES - 1):
            imgs.append(r.GetImage().copy())
    return cache, r, imgs, ms


cache, r, imgs, ms = run(True)
t = torch.tensor(ms, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
ok = True
# (2) replicas identical
w = torch.from_numpy(cache.get_params(N.EMA)).cuda()
gl = [torch.empty_like(w) for _ in range(world)]
dist.all_gather(gl, w)
same_params = all(torch.equal(gl[0], x) for x in gl)
ok = ok and same_params
# (1) stitch frame 0 and the last frame on rank 0
first = torch.from_numpy(imgs[0]).cuda(); last = torch.from_numpy(imgs[1]).cuda()
for img in (first, last):
    parts = [torch.empty_like(img) for _ in range(world)]
    dist.all_gather(parts, img)
    if rank == 0:
        full = torch.zeros_like(img)
        for (b, e), p in zip(strips, parts):
            full[:, b:e] = p[:, b:e]
        img.copy_(full)
if rank == 0:
    _, _, ref_imgs, ref_ms = run(False)                  # the whole frame on one GPU
    a, b = first.cpu().numpy(), ref_imgs[0]
    frame0_equal = bool(np.array_equal(a, b, equal_nan=True))
    lastn = last.cpu().numpy()
    finite = float(np.isfinite(lastn).mean())
    differs = bool(np.any(lastn != a))
    ok = ok and frame0_equal and finite > 0.9999 and differs
    print(json.dumps({"world": world, "resolution": [W, H], "strips": strips, "train_records_per_rank_and_step": tile_app.train_batch_size, "frame0_stitched_equals_single_gpu": frame0_equal, "replica_parameters_bit_identical": same_params,
                      "last_frame_finite_fraction": finite, "ms_per_frame_tiles_max_over_ranks": round(float(t[2:].mean()), 4),
                      "ms_per_frame_single_gpu": round(float(np.mean(ref_ms[2:])), 4), "loss": cache.GetLoss()}))
flag = torch.tensor([1 if ok else 0], device="cuda"); dist.broadcast(flag, 0)
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
