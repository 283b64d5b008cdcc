"""Training-step timing on one B200 (CUDA events, 4 record sets rotated): the fused kernel (NRCHPM_TRAIN_FUSED / NRCHPM_TRAIN_TPR knobs of
csrc/nrc.cu) against the three-kernel path, at the reference's 2^14-record step and at large batches.  Development tool."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_records
from nrc_hpm_renderer_b200 import AppConfig
from nrc_hpm_renderer_b200.nrc import NeuralRadianceCache

def run(fused, tpr, pos=0, depth=6, sizes=(1 << 14, 1 << 18, 1 << 20)):
    os.environ["NRCHPM_TRAIN_FUSED"] = str(fused); os.environ["NRCHPM_TRAIN_TPR"] = str(tpr)
    app = AppConfig.default(); app.pos_enc_id, app.nn_depth = pos, depth
    nrc = NeuralRadianceCache(app)
    st = torch.cuda.current_stream(); sp = st.cuda_stream
    rng = np.random.default_rng(1337)
    nmax = max(sizes)
    tin = [torch.from_numpy(synth_records(rng, nmax)).cuda() for _ in range(4)]
    tgt = [torch.from_numpy((rng.random((nmax, 3), dtype=np.float32) * 2).astype(np.float32)).cuda() for _ in range(4)]
    res = {"fused": fused, "tpr": tpr, "pos": pos, "depth": depth}
    losses = []
    for i in range(8):
        nrc.training_step(tin[i % 4][:16384], tgt[i % 4][:16384], 16384, True, sp); losses.append(nrc.GetLoss())
    res["loss8"] = losses[-1]; res["loss1"] = losses[0]
    for n in sizes:
        iters = max(5, min(200, (1 << 22) // n))
        for i in range(3): nrc.training_step(tin[i % 4][:n], tgt[i % 4][:n], n, True, sp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for i in range(iters): nrc.training_step(tin[i % 4][:n], tgt[i % 4][:n], n, True, sp)
        e1.record(st); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / iters
        res[f"step_us_{n}"] = round(t * 1e3, 1); res[f"samples_per_s_{n}"] = n / t * 1e3
    nrc.Destroy()
    return res

if __name__ == "__main__":
    for fused, tpr in ((0, 2), (1, 2), (1, 4)):
        print(json.dumps(run(fused, tpr)), flush=True)
    for fused, tpr in ((0, 2), (1, 2), (1, 4)):
        print(json.dumps(run(fused, tpr, pos=2, depth=5)), flush=True)
