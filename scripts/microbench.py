"""BASELINE config 3: synthetic NRC microbench -- 2^18 .. 2^22 query records, 64-wide MLP with 5 and 6 hidden
layers, (i) HashGrid16x2 + OneBlob4 (reference default) and (ii) TriangleWave12 + OneBlob4 (no encoding parameters: isolates the
MLP / tensor pipe); inference and training throughput against the measured tensor peak.  Prints one JSON line per case.
CUDA events on the launch stream, 4 record sets rotated (inputs larger than L2 for the larger sizes), 3 warm-up rounds."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import synth_records, peaks
from nrc_hpm_renderer_b200 import AppConfig
from nrc_hpm_renderer_b200.nrc import NeuralRadianceCache

pk = peaks()
st = torch.cuda.current_stream(); sp = st.cuda_stream
rng = np.random.default_rng(1337)
NMAX = 1 << 22
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


for pos, pos_name in ((0, "HashGrid16x2"), (2, "TriangleWave12")):
    for depth in (5, 6):
        app = AppConfig.default(); app.pos_enc_id, app.dir_enc_id, app.nn_depth = pos, 0, depth
        c = NeuralRadianceCache(app)
        in_w = c.input_width
        flop = 2 * (in_w * 64 + (depth - 1) * 64 * 64 + 64 * 3)
        for i in range(4): c.training_step(d_in[i][:16384], d_tgt[0][:16384], 16384, True, sp)     # non-trivial weights / EMA
        for log2n in (18, 19, 20, 21, 22):
            n = 1 << log2n
            ms_i = timed(lambda i: c.inference(d_in[i % 4], d_out, n, True, sp), 10)
            ms_t = timed(lambda i: c.training_step(d_in[i % 4][:n], d_tgt[i % 2][:n], n, True, sp), 5)
            print(json.dumps({"encoding": pos_name + "+OneBlob4", "hidden_layers": depth, "input_width": in_w, "records": n,
                              "inference_ms": round(ms_i, 4), "inference_queries_per_s": n / ms_i * 1e3, "inference_tflops": flop * n / ms_i / 1e9,
                              "inference_frac_of_sustained_tensor_peak": flop * n / ms_i / 1e9 / pk["tflops_sustained"],
                              "training_ms": round(ms_t, 4), "training_samples_per_s": n / ms_t * 1e3, "training_tflops": 3 * flop * n / ms_t / 1e9,
                              "training_frac_of_sustained_tensor_peak": 3 * flop * n / ms_t / 1e9 / pk["tflops_sustained"]}), flush=True)
        c.Destroy()
