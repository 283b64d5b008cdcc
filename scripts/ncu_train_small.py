"""Development tool: a few reference-sized (2^14) training steps for ncu captures of the training kernels / optimizer."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_records
from nrc_hpm_renderer_b200 import AppConfig
from nrc_hpm_renderer_b200.nrc import NeuralRadianceCache
n = 1 << 14
rng = np.random.default_rng(1)
c = NeuralRadianceCache(AppConfig.default())
d_in = [torch.from_numpy(synth_records(rng, n)).cuda() for _ in range(4)]
d_tgt = [torch.from_numpy((rng.random((n, 3), dtype=np.float32) * 2).astype(np.float32)).cuda() for _ in range(4)]
for i in range(24): c.training_step(d_in[i % 4], d_tgt[i % 4], n, True)
torch.cuda.synchronize()
