"""Development tool: two training steps at batch 2^20 for an ncu launch list."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_records
from nrc_hpm_renderer_b200 import AppConfig
from nrc_hpm_renderer_b200.nrc import NeuralRadianceCache
n = 1 << 20
rng = np.random.default_rng(1)
c = NeuralRadianceCache(AppConfig.default())
d_in = torch.from_numpy(synth_records(rng, n)).cuda(); d_tgt = torch.from_numpy((rng.random((n, 3), dtype=np.float32) * 2).astype(np.float32)).cuda()
for i in range(3): c.training_step(d_in, d_tgt, n, True)
torch.cuda.synchronize()
