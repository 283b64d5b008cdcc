"""Development tool (GPU): e2e step time of nrc_infer_and_train_host under the experiment knobs."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_records, N_INFER, TRAIN_BATCH, TRAIN_BATCHES
from nrc_hpm_renderer_b200 import AppConfig
from nrc_hpm_renderer_b200.nrc import NeuralRadianceCache
nrc = NeuralRadianceCache(AppConfig.default())
rng = np.random.default_rng(1)
n_train = TRAIN_BATCH * TRAIN_BATCHES
pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
h_in = [pin(synth_records(rng, N_INFER)) for _ in range(2)]
h_out = pin(np.zeros((N_INFER, 3), np.float32))
h_tin = [pin(synth_records(rng, n_train)) for _ in range(2)]
h_tgt = [pin((rng.random((n_train, 3), dtype=np.float32) * 2).astype(np.float32)) for _ in range(2)]
for i in range(5): nrc.infer_and_train_host(h_in[i % 2], h_out, h_tin[i % 2], h_tgt[i % 2], TRAIN_BATCH, True)
torch.cuda.synchronize(); t0 = time.perf_counter(); K = 50
for i in range(K): nrc.infer_and_train_host(h_in[i % 2], h_out, h_tin[i % 2], h_tgt[i % 2], TRAIN_BATCH, True)
torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / K * 1e3
# inference only / training only for the breakdown
t0 = time.perf_counter()
for i in range(K): nrc.inference_host(h_in[i % 2], out=h_out)
torch.cuda.synchronize(); ms_inf = (time.perf_counter() - t0) / K * 1e3
t0 = time.perf_counter()
for i in range(K):
    for b in range(TRAIN_BATCHES): nrc.training_step_host(h_tin[i % 2][b * TRAIN_BATCH:(b + 1) * TRAIN_BATCH], h_tgt[i % 2][b * TRAIN_BATCH:(b + 1) * TRAIN_BATCH])
torch.cuda.synchronize(); ms_tr = (time.perf_counter() - t0) / K * 1e3
print(json.dumps({"chunk_tiles": os.environ.get("NRCHPM_E2E_CHUNK_TILES", "4"), "e2e_ms": round(ms, 4), "inference_host_only_ms": round(ms_inf, 4), "train_host_only_ms": round(ms_tr, 4)}))
