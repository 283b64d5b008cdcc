"""Development tool (GPU): why do compact and all-records inference differ with identical parameters?"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import oracle as O
import test_gpu_tracker as T
from nrc_hpm_renderer_b200 import AppConfig, renderer as R
from nrc_hpm_renderer_b200 import nrc as N
O.build()
W, H = 128, 64
caches, rs = [], []
for compact in (True, False):
    app = AppConfig.default(); app.log2_train_batch_size, app.train_batch_count, app.log2_infer_batch_size = 9, 2, 12
    nrc = N.NeuralRadianceCache(app)
    r, *_ = T.setup(0, W, H, O, train_pixels=1024, compact=compact, nrc=nrc)
    for f in range(3):
        r.Render(True, T.FR + np.float32(0.05 * f))
    caches.append(nrc); rs.append(r)
caches[1].set_params(caches[0].get_params(N.MASTER)); caches[1].set_ema(caches[0].get_params(N.EMA))
for which, name in ((N.MASTER, "master"), (N.WORKING, "working"), (N.EMA, "ema")):
    a, b = caches[0].get_params(which), caches[1].get_params(which)
    print(json.dumps({"params": name, "equal": bool(np.array_equal(a, b)), "n_diff": int((a != b).sum()), "max": float(np.abs(a - b).max())}))
fr = T.FR + np.float32(0.3)
for r in rs: r.Render(False, fr)
inp = [r.read(R.BUF_INFER_INPUT).reshape(-1, 5) for r in rs]
out = [r.read(R.BUF_INFER_OUTPUT).reshape(-1, 3) for r in rs]
info = [r.read(R.BUF_PRIMARY_INFO) for r in rs]           # pixel order y*W+x
img = [r.GetImage() for r in rs]
print(json.dumps({"inputs_equal": bool(np.array_equal(inp[0], inp[1], equal_nan=True)), "info_equal": bool(np.array_equal(info[0], info[1]))}))
sc = info[0].reshape(H, W).T.reshape(-1) == 1.0            # record order x*H+y
d = np.abs(out[0] - out[1]); d[~np.isfinite(d)] = 0
print(json.dumps({"scattered": int(sc.sum()), "out_diff_scattered": int((d[sc] != 0).any(1).sum()), "max": float(d[sc].max()), "nan_rows": [int(np.isnan(o[sc]).any(1).sum()) for o in out],
                  "nan_inputs": int(np.isnan(inp[0][sc]).any(1).sum())}))
# independent evaluation through the batch API of each cache
rec = np.ascontiguousarray(inp[0])
ev = [c.inference_host(rec, use_ema=True) for c in caches]
print(json.dumps({"batch_api_equal_between_caches": bool(np.array_equal(ev[0], ev[1], equal_nan=True)),
                  "compact_vs_batch": float(np.nan_to_num(np.abs(out[0] - ev[0])[sc]).max()), "all_vs_batch": float(np.nan_to_num(np.abs(out[1] - ev[1])[sc]).max())}))
bad = np.where(sc & (d != 0).any(1))[0][:5]
for i in bad: print(int(i), inp[0][i].tolist(), out[0][i].tolist(), out[1][i].tolist(), ev[0][i].tolist())
di = np.abs(img[0] - img[1]); di[~np.isfinite(di)] = 0
print(json.dumps({"img_diff_pixels": int((di != 0).any(2).sum()), "img_max": float(di.max())}))
