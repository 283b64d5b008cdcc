"""Inference-launch timing on one B200 (CUDA events, 4 record sets rotated through > L2): kernel variants selected by NRCHPM_INFER_WS.
Development tool."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_records, N_INFER, TRAIN_BATCH
from nrc_hpm_renderer_b200 import AppConfig
from nrc_hpm_renderer_b200.nrc import NeuralRadianceCache

def run(ws, ref=None, sizes=(N_INFER, 1 << 18, 1 << 20, 1 << 22), pos=0, depth=6):
    os.environ["NRCHPM_INFER_WS"] = str(ws)
    app = AppConfig.default(); app.pos_enc_id, app.nn_depth = pos, depth
    nrc = NeuralRadianceCache(app)
    st = torch.cuda.current_stream(); sp = st.cuda_stream
    rng = np.random.default_rng(1337)
    nmax = max(sizes)
    d_in = [torch.from_numpy(synth_records(rng, nmax)).cuda() for _ in range(4)]
    d_out = torch.empty((nmax, 3), dtype=torch.float32, device="cuda")
    tgt = torch.from_numpy((rng.random((TRAIN_BATCH, 3), dtype=np.float32) * 2).astype(np.float32)).cuda()
    for i in range(4): nrc.training_step(d_in[i][:TRAIN_BATCH], tgt, TRAIN_BATCH, True, sp)
    nrc.inference(d_in[0], d_out, N_INFER, True, sp); torch.cuda.synchronize()
    out0 = d_out[:N_INFER].cpu().numpy().copy()
    res = {"ws": ws, "pos": pos, "depth": depth}
    for n in sizes:
        for i in range(3): nrc.inference(d_in[i % 4], d_out, n, True, sp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20
        e0.record(st)
        for i in range(iters): nrc.inference(d_in[i % 4], d_out, n, True, sp)
        e1.record(st); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / iters
        res[f"ms_{n}"] = round(t, 4); res[f"qps_{n}"] = n / t * 1e3
    if ref is not None:
        res["bit_identical_to_ws0"] = bool(np.array_equal(out0, ref))
        res["max_abs_diff"] = float(np.nanmax(np.abs(out0 - ref)))
    nrc.Destroy()
    return res, out0

if __name__ == "__main__":
    ref = None
    for ws in (0, 1, 2):
        r, out = run(ws, ref)
        if ref is None: ref = out
        print(json.dumps(r), flush=True)
    for ws in (0, 1, 2):
        r, _ = run(ws, None, pos=2, depth=5)
        print(json.dumps(r), flush=True)
