"""Generate tests/golden/tcnn_<cfg>.npz on a GPU box by running the REFERENCE's own tiny-cuda-nn
(oracle/_ref/tcnn_oracle, built from /root/reference/tiny-cuda-nn by oracle/tcnn_ref/Makefile).

    gpurun -- python tests/golden/make_tcnn_golden.py          # writes gpurun_out/tcnn_*.npz
    cp gpurun_out/tcnn_*.npz tests/golden/

The fixtures pin both the CPU oracle (oracle/nrc_oracle.cpp) and the CUDA product against the
reference arithmetic: initial parameters (bit-exact), network_input, inference outputs, step-0
gradients, loss curve, parameters/EMA after the last step.  Grid tensors are stored sparsely (only the
entries touched by step 0 plus a strided sample) to keep the fixtures small.
"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
BIN = os.path.join(ROOT, "oracle", "_ref", "tcnn_oracle")
SKY_HALF = np.array([62.317, 42.295, 76.707], dtype=np.float32) / 2  # skySize/2

CONFIGS = [  # name, pos, dir, depth, n_infer, batch, steps, n_neurons
    ("hash_ob_d6", 0, 0, 6, 1024, 256, 16, 64),
    ("tri_ob_d5", 2, 0, 5, 1024, 256, 16, 64),
    ("hash_tri_d3", 0, 2, 3, 512, 256, 4, 64),
    ("id_id_d2", 1, 1, 2, 512, 256, 4, 64),
    ("freq_ob_d4", 3, 0, 4, 512, 256, 4, 64),
    # nnWidth = 128 (reference src/AppConfig.cpp:169; tcnn instantiates FullyFusedMLP<__half, 128>, src/network.cu:117-118)
    ("hash_ob_d6_w128", 0, 0, 6, 1024, 256, 16, 128),
    ("tri_ob_d5_w128", 2, 0, 5, 1024, 256, 16, 128),
]


def make_records(rng, n):
    rec = np.empty((n, 5), dtype=np.float32)
    rec[:, :3] = rng.random((n, 3), dtype=np.float32)
    rec[: n // 2, :3] += SKY_HALF                      # reference-style positions (Q4)
    rec[:, 3] = rng.random(n, dtype=np.float32) * 2 - 0.5
    rec[:, 4] = rng.random(n, dtype=np.float32)
    rec[rng.random(n) < 0.1, 4] = np.nan               # acos(>1) in the reference's phi (Q5)
    return rec


def main():
    out_root = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_root, exist_ok=True)
    only = set(sys.argv[1:])
    for name, pos, dr, depth, n_infer, batch, steps, width in CONFIGS:
        if only and name not in only:
            continue
        rng = np.random.default_rng(1337 + pos * 10 + dr)
        work = os.path.join(out_root, "tcnn_dump_" + name)
        os.makedirs(work, exist_ok=True)
        infer_in = make_records(rng, n_infer)
        train_in = make_records(rng, steps * batch)
        train_tgt = (rng.random((steps * batch, 3), dtype=np.float32) * 2).astype(np.float32)
        infer_in.tofile(work + "/infer_in.f32"); train_in.tofile(work + "/train_in.f32"); train_tgt.tofile(work + "/train_tgt.f32")
        cmd = [BIN, "dump", f"out={work}", f"pos={pos}", f"dir={dr}", f"depth={depth}", f"width={width}", f"n_infer={n_infer}", f"batch={batch}",
               f"steps={steps}", f"infer_in={work}/infer_in.f32", f"train_in={work}/train_in.f32", f"train_tgt={work}/train_tgt.f32"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        print(res.stdout, res.stderr, file=sys.stderr)
        if res.returncode != 0:
            raise SystemExit(f"tcnn_oracle failed for {name}")
        meta = {}
        for line in res.stdout.splitlines():
            if line.startswith("{"):
                meta.update(json.loads(line))
        P, inw = meta["n_params"], meta["padded_input"]
        n_mlp = width * inw + (depth - 1) * width * width + 16 * width
        f32 = lambda f: np.fromfile(f"{work}/{f}", dtype=np.float32)
        f16 = lambda f: np.fromfile(f"{work}/{f}", dtype=np.float16)
        grad0 = f16("grad_step0.f16")
        touched = np.nonzero(grad0[n_mlp:])[0].astype(np.int64) + n_mlp
        sample = np.arange(n_mlp, P, 4099, dtype=np.int64)
        grid_idx = np.union1d(touched, sample)
        params_init = f32("params_init.f32")
        net_in = f16("network_input.f16")
        net_in = net_in.reshape(n_infer, inw) if meta["network_input_layout"] == "AoS" else net_in.reshape(inw, n_infer).T
        pick = lambda a: a[grid_idx]
        np.savez_compressed(
            os.path.join(out_root, f"tcnn_{name}.npz"),
            pos=pos, dir=dr, depth=depth, width=width, n_infer=n_infer, batch=batch, steps=steps, n_params=P, n_mlp=n_mlp, padded_input=inw,
            infer_in=infer_in, train_in=train_in, train_tgt=train_tgt,
            grid_idx=grid_idx,
            params_init_mlp=params_init[:n_mlp], params_init_grid=pick(params_init),
            params_init_sum=np.float64(params_init.astype(np.float64).sum()), params_init_abs_sum=np.float64(np.abs(params_init.astype(np.float64)).sum()),
            network_input=np.ascontiguousarray(net_in),
            infer_ema_step0=f32("infer_ema_step0.f32").reshape(n_infer, 3),
            infer_working_step0=f32("infer_working_step0.f32").reshape(n_infer, 3),
            output_step0=f16("output_step0.f16").reshape(batch, 16), dL_doutput_step0=f16("dL_doutput_step0.f16").reshape(batch, 16),
            grad0_mlp=grad0[:n_mlp], grad0_grid=pick(grad0),
            params_step1_mlp=f32("params_step1.f32")[:n_mlp], params_step1_grid=pick(f32("params_step1.f32")),
            ema_step1_mlp=f16("ema_step1.f16")[:n_mlp], ema_step1_grid=pick(f16("ema_step1.f16")),
            losses=f32("losses.f32"),
            params_final_mlp=f32("params_final.f32")[:n_mlp], params_final_grid=pick(f32("params_final.f32")),
            ema_final_mlp=f16("ema_final.f16")[:n_mlp], ema_final_grid=pick(f16("ema_final.f16")),
            infer_ema_final=f32("infer_ema_final.f32").reshape(n_infer, 3),
        )
        for f in os.listdir(work):
            os.remove(os.path.join(work, f))
        os.rmdir(work)
        print("wrote", name, "P=", P, "touched=", len(touched), "losses", f32 if False else "", file=sys.stderr)


if __name__ == "__main__":
    main()
