"""BASELINE config 3, reference column: the reference's own tiny-cuda-nn (oracle/_ref/tcnn_oracle, unmodified sources built for sm_100a) on
the same synthetic batches as scripts/microbench.py -- inference-only and training-only throughput at 2^18 .. 2^22 records
(protocol of tiny-cuda-nn/benchmarks/image/bench_ours.cu:188-331: CUDA events around back-to-back iterations after warm-up).
One JSON line per case -> profiles/r02_microbench_c3_tcnn.jsonl."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "tcnn_oracle")
for pos, pos_name in ((0, "HashGrid16x2"), (2, "TriangleWave12")):
    for depth in (5, 6):
        for log2n in (14, 18, 19, 20, 21, 22):
            n = 1 << log2n
            row = {"impl": "tcnn", "encoding": pos_name + "+OneBlob4", "hidden_layers": depth, "records": n}
            for mode in ("inference", "training"):
                if mode == "inference" and log2n == 14:
                    continue
                frames = max(10, min(200, (1 << 24) // n))
                cmd = [BIN, "bench", f"n_infer={n}", f"infer_batch={n}", f"batch={n}", "batches=1", f"frames={frames}", "warmup=5", f"pos={pos}", "dir=0", f"depth={depth}", "sets=4",
                       f"infer={1 if mode == 'inference' else 0}", f"train={1 if mode == 'training' else 0}"]
                res = subprocess.run(cmd, capture_output=True, text=True)
                line = [l for l in res.stdout.splitlines() if l.startswith("{")]
                if res.returncode != 0 or not line:
                    row[mode + "_error"] = (res.stderr.strip().splitlines() or ["?"])[-1][:200]
                    continue
                j = json.loads(line[-1])
                row[mode + "_ms"] = j["ms_per_frame"]
                row[mode + ("_queries_per_s" if mode == "inference" else "_samples_per_s")] = n / j["ms_per_frame"] * 1e3
            print(json.dumps(row), flush=True)
