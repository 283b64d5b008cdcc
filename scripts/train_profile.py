"""Development tool (GPU box): phase breakdown of the fused training kernel (NRCHPM_TRAIN_PROF=1: clock64 stamps of thread 0 of every CTA)
and the device-side timeline of consecutive training steps (fused step, optimizer, EMA pass)."""
import ctypes as C, os, sys, json
os.environ["NRCHPM_TRAIN_PROF"] = "1"
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_records
from nrc_hpm_renderer_b200 import AppConfig, _lib
from nrc_hpm_renderer_b200.nrc import NeuralRadianceCache
names = ["setup", "encode", "sync", "mma0", "fwd_layers", "loss+first_bwd_mma", "bwd_layers", "scatter", "dw_readout", "loss_reduce"]
for tpr in [int(v) for v in (sys.argv[1:] or ["2", "4"])]:
    os.environ["NRCHPM_TRAIN_TPR"] = str(tpr)
    n = 1 << 14
    rng = np.random.default_rng(1)
    c = NeuralRadianceCache(AppConfig.default())
    d_in = [torch.from_numpy(synth_records(rng, n)).cuda() for _ in range(4)]
    d_tgt = [torch.from_numpy((rng.random((n, 3), dtype=np.float32) * 2).astype(np.float32)).cuda() for _ in range(4)]
    acc = []
    for i in range(12):
        c.training_step(d_in[i % 4], d_tgt[i % 4], n, True)
        if i >= 4:
            buf = np.zeros((148, 16), np.int64)
            _lib.lib().nrc_debug_train_profile(c._h, buf.ctypes.data_as(C.c_void_p), 148)
            acc.append(np.diff(buf[:128, :11], axis=1).mean(0))
    m = np.mean(acc, 0) / 1.965e3       # us at 1965 MHz
    print(json.dumps({"tpr": tpr, "phases_us": {nm: round(float(v), 2) for nm, v in zip(names, m)}, "total_us": round(float(m.sum()), 2)}))
    tl = np.zeros(512, np.uint64)
    _lib.lib().nrc_debug_timeline(c._h, tl.ctypes.data_as(C.c_void_p), 256)       # reset
    for i in range(8):
        c.training_step(d_in[i % 4], d_tgt[i % 4], n, True)
    k = _lib.lib().nrc_debug_timeline(c._h, tl.ctypes.data_as(C.c_void_p), 256)
    t = tl[:2 * k].astype(np.float64).reshape(k, 2)
    t = (t - t[0, 0]) / 1e3
    rows = [{"kernel": ("fused", "adam", "ema")[j % 3], "begin_us": round(float(b), 1), "end_us": round(float(e), 1)} for j, (b, e) in enumerate(t)]
    print(json.dumps({"timeline": rows[:15]}))
    c.Destroy()
