"""Development tool (torchrun, N >= 2): device-side timeline of the data-parallel InferAndTrain schedule on rank 0 (NRCHPM_TRAIN_PROF=1).
Launch order per frame with head h > 0: [infer head] then per batch: fused, gather (reduce-scatter), infer chunk, adam (own slice), ema, publish (weight all-gather) -- slot order, the EMA pass RUNS after the publish."""
import ctypes as C, json, os, sys
os.environ["NRCHPM_TRAIN_PROF"] = "1"
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from bench import synth_records, N_INFER, TRAIN_BATCH, TRAIN_BATCHES
from nrc_hpm_renderer_b200 import AppConfig, _lib
from nrc_hpm_renderer_b200.nrc import NeuralRadianceCache
from nrc_hpm_renderer_b200.parallel import PeerGradientExchange
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nrc = NeuralRadianceCache(AppConfig.default())
PeerGradientExchange(nrc, world)
rng = np.random.default_rng(1337 + rank)
d_in = torch.from_numpy(synth_records(rng, N_INFER)).cuda(); d_out = torch.empty((N_INFER, 3), dtype=torch.float32, device="cuda")
n_train = TRAIN_BATCH * TRAIN_BATCHES
d_tin = torch.from_numpy(synth_records(rng, n_train)).cuda(); d_tgt = torch.from_numpy((rng.random((n_train, 3), dtype=np.float32) * 2).astype(np.float32)).cuda()
sp = torch.cuda.current_stream().cuda_stream
nrc.Init(N_INFER, d_in, d_out, d_tin, d_tgt, None, None, sp)
tl = np.zeros(512, np.uint64)
for i in range(6):
    nrc.InferAndTrain(None, True)
torch.cuda.synchronize(); dist.barrier()
_lib.lib().nrc_debug_timeline(nrc._h, tl.ctypes.data_as(C.c_void_p), 256)
for i in range(2):
    nrc.InferAndTrain(None, True)
torch.cuda.synchronize(); dist.barrier()
k = _lib.lib().nrc_debug_timeline(nrc._h, tl.ctypes.data_as(C.c_void_p), 256)
if rank == 0:
    t = tl[:2 * k].astype(np.float64).reshape(k, 2); t = (t - t[:, 0].min()) / 1e3
    fused = os.environ.get("NRCHPM_PEER_FUSED", "1") != "0"
    overlap = os.environ.get("NRCHPM_OVERLAP", "1") != "0"
    if fused and overlap: per_batch, lead = ["fused", "infer_chunk", "peer_adam", "ema"], []
    elif fused: per_batch, lead = ["fused", "peer_adam", "ema"], ["inference"]
    elif overlap: per_batch, lead = ["fused", "gather", "infer_chunk", "adam", "ema", "publish"], []
    else: per_batch, lead = ["fused", "gather", "adam", "ema", "publish"], ["inference"]
    if overlap and float(os.environ.get("NRCHPM_OVERLAP_HEAD", "0")) > 0: lead = ["infer_head"]
    names = lead + per_batch * TRAIN_BATCHES
    per = len(names)
    for j, (b, e) in enumerate(t):
        print(f"{names[j % per]:12s} {b:9.1f} {e:9.1f}  ({e - b:7.1f} us)")
dist.destroy_process_group()
