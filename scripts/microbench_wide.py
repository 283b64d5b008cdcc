"""128-neuron network (nnWidth = 128, reference src/AppConfig.cpp:169): inference on a 1080p frame of records and training at 2^14 /
2^18 / 2^20 records, this library (csrc/nrc_wide_kernels.cuh) next to the reference's own tiny-cuda-nn (oracle/_ref/tcnn_oracle,
FullyFusedMLP<__half, 128>), same synthetic records, CUDA events after warm-up.  One JSON line per case."""
import json, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import synth_records, peaks, N_INFER
from nrc_hpm_renderer_b200 import AppConfig
from nrc_hpm_renderer_b200.nrc import NeuralRadianceCache

BIN = os.path.join(ROOT, "oracle", "_ref", "tcnn_oracle")
pk = peaks()
st = torch.cuda.current_stream(); sp = st.cuda_stream
rng = np.random.default_rng(1337)
NMAX = 1 << 21
d_in = [torch.from_numpy(synth_records(rng, NMAX)).cuda() for _ in range(4)]
d_tgt = [torch.from_numpy((rng.random((NMAX, 3), dtype=np.float32) * 2).astype(np.float32)).cuda() for _ in range(2)]
d_out = torch.empty((NMAX, 3), dtype=torch.float32, device="cuda")


def timed(fn, iters):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for i in range(iters): fn(i)
    e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def tcnn(n, pos, depth, mode):
    if not os.path.exists(BIN):
        return None
    frames = max(10, min(200, (1 << 24) // n))
    cmd = [BIN, "bench", f"n_infer={n}", f"infer_batch={n}", f"batch={n}", "batches=1", f"frames={frames}", "warmup=5", f"pos={pos}", "dir=0", f"depth={depth}", "width=128", "sets=4",
           f"infer={1 if mode == 'inference' else 0}", f"train={1 if mode == 'training' else 0}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    line = [l for l in res.stdout.splitlines() if l.startswith("{")]
    return json.loads(line[-1])["ms_per_frame"] if res.returncode == 0 and line else None


for pos, pos_name, depth in ((0, "HashGrid16x2", 6), (2, "TriangleWave12", 5)):
    app = AppConfig.default(); app.pos_enc_id, app.dir_enc_id, app.nn_depth, app.nn_width = pos, 0, depth, 128
    c = NeuralRadianceCache(app)
    in_w = c.input_width
    flop = 2 * (in_w * 128 + (depth - 1) * 128 * 128 + 128 * 3)
    for i in range(4): c.training_step(d_in[i][:16384], d_tgt[0][:16384], 16384, True, sp)
    n = N_INFER
    ms = timed(lambda i: c.inference(d_in[i % 4], d_out, n, True, sp), 10)
    ref = tcnn(n, pos, depth, "inference")
    print(json.dumps({"n_neurons": 128, "encoding": pos_name + "+OneBlob4", "hidden_layers": depth, "what": "inference", "records": n, "ms": round(ms, 4), "queries_per_s": n / ms * 1e3,
                      "tflops": flop * n / ms / 1e9, "frac_of_sustained_tensor_peak": flop * n / ms / 1e9 / pk["tflops_sustained"], "tcnn_ms": ref, "speedup_vs_tcnn": (ref / ms) if ref else None}), flush=True)
    for log2n in (14, 18, 20):
        n = 1 << log2n
        ms = timed(lambda i: c.training_step(d_in[i % 4][:n], d_tgt[i % 2][:n], n, True, sp), 20 if log2n == 14 else 5)
        ref = tcnn(n, pos, depth, "training")
        print(json.dumps({"n_neurons": 128, "encoding": pos_name + "+OneBlob4", "hidden_layers": depth, "what": "training", "records": n, "ms": round(ms, 4), "samples_per_s": n / ms * 1e3,
                          "tflops": 3 * flop * n / ms / 1e9, "frac_of_sustained_tensor_peak": 3 * flop * n / ms / 1e9 / pk["tflops_sustained"], "tcnn_ms": ref, "speedup_vs_tcnn": (ref / ms) if ref else None}), flush=True)
    c.Destroy()
