"""Development tool: a few inference launches of the bench workload, for `ncu -k regex:nrc_forward -s 2 -c 1` captures."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_records, N_INFER
from nrc_hpm_renderer_b200 import AppConfig
from nrc_hpm_renderer_b200.nrc import NeuralRadianceCache

nrc = NeuralRadianceCache(AppConfig.default())
sp = torch.cuda.current_stream().cuda_stream
rng = np.random.default_rng(1337)
d_in = [torch.from_numpy(synth_records(rng, N_INFER)).cuda() for _ in range(2)]
d_out = torch.empty((N_INFER, 3), dtype=torch.float32, device="cuda")
for i in range(4):
    nrc.inference(d_in[i % 2], d_out, N_INFER, False, sp)
torch.cuda.synchronize()
