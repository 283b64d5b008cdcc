"""Development tool (GPU box): the CUDA path's 1000-step loss curves for every data order of tests/golden/make_tcnn_loss_curve.py's
multi-seed fixture -> gpurun_out/ours_loss1000_<cfg>_seeds.npz (used to state the tolerance of tests/test_gpu_loss_curve.py)."""
import importlib.util, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nrc_hpm_renderer_b200 import AppConfig
from nrc_hpm_renderer_b200.nrc import NeuralRadianceCache
spec = importlib.util.spec_from_file_location("g", os.path.join(ROOT, "tests", "golden", "make_tcnn_loss_curve.py"))
g = importlib.util.module_from_spec(spec); spec.loader.exec_module(g)
for name, pos, dr, depth, batch, steps, lr in g.CONFIGS:
    seed = 4242 + pos * 10 + dr
    curves = []
    for o in g.SEED_OFFSETS:
        tin, tgt, held = g.training_data(seed + o, batch, steps)
        app = AppConfig.default(); app.pos_enc_id, app.dir_enc_id, app.nn_depth, app.learning_rate = pos, dr, depth, lr
        c = NeuralRadianceCache(app)
        d_in, d_tgt = torch.from_numpy(tin).cuda(), torch.from_numpy(tgt).cuda()
        L = np.empty(steps, np.float32)
        for s in range(steps):
            c.training_step(d_in[s * batch:(s + 1) * batch], d_tgt[s * batch:(s + 1) * batch], batch, True)
            L[s] = c.GetLoss()
        curves.append(L); c.Destroy()
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"ours_loss1000_{name}_seeds.npz"), losses=np.stack(curves))
    print(name, "done", flush=True)
