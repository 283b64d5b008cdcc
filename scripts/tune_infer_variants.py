"""Development tool: inference launch time (2 073 600 records, hash grid, 64 x 6) of the product library and of every variant under
nrc_hpm_renderer_b200/variants/ (each in its own process: NRCHPM_LIB)."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, %r)
import torch
from bench import synth_records, N_INFER, TRAIN_BATCH
from nrc_hpm_renderer_b200 import AppConfig
from nrc_hpm_renderer_b200.nrc import NeuralRadianceCache
nrc = NeuralRadianceCache(AppConfig.default())
st = torch.cuda.current_stream(); sp = st.cuda_stream
rng = np.random.default_rng(1337)
d_in = [torch.from_numpy(synth_records(rng, N_INFER)).cuda() for _ in range(4)]
d_out = torch.empty((N_INFER, 3), dtype=torch.float32, device="cuda")
nrc.set_ema(nrc.get_params(0))
for i in range(4): nrc.inference(d_in[i], d_out, N_INFER, True, sp)
torch.cuda.synchronize()
crc = int(np.bitwise_xor.reduce(d_out.cpu().numpy().view(np.uint32).ravel()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for i in range(40): nrc.inference(d_in[i %% 4], d_out, N_INFER, True, sp)
e1.record(st); torch.cuda.synchronize()
print(json.dumps({"lib": os.environ.get("NRCHPM_LIB", "product"), "ws": os.environ.get("NRCHPM_INFER_WS", "1"), "infer_ms": round(e0.elapsed_time(e1) / 40, 4), "out_xor": crc}))
''' % ROOT
runs = [({"NRCHPM_INFER_WS": "0"}, None), ({}, None)] + [({}, v) for v in sorted(glob.glob(os.path.join(ROOT, "nrc_hpm_renderer_b200", "variants", "*.so")))]
for env_extra, lib in runs:
    env = dict(os.environ); env.update(env_extra)
    if lib: env["NRCHPM_LIB"] = lib
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    print(r.stdout.strip() or r.stderr.strip()[-300:], flush=True)
