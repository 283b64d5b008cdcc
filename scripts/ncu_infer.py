"""Development tool: a few full-frame inference launches for ncu captures of the dominant kernel."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_records, N_INFER
from nrc_hpm_renderer_b200 import AppConfig
from nrc_hpm_renderer_b200.nrc import NeuralRadianceCache
rng = np.random.default_rng(1)
c = NeuralRadianceCache(AppConfig.default())
c.set_ema(c.get_params(0))
d_in = [torch.from_numpy(synth_records(rng, N_INFER)).cuda() for _ in range(4)]
d_out = torch.empty((N_INFER, 3), dtype=torch.float32, device="cuda")
for i in range(6): c.inference(d_in[i % 4], d_out, N_INFER, True)
torch.cuda.synchronize()
