"""Generate tests/golden/tcnn_loss1000_<cfg>.npz on a GPU box: the REFERENCE's own tiny-cuda-nn (oracle/_ref/tcnn_oracle, built
unmodified from /root/reference/tiny-cuda-nn by oracle/tcnn_ref/Makefile) trains for 1000 steps on a learnable synthetic
radiance field, driven like en::NeuralRadianceCache::Train (reference src/NeuralRadianceCache.cu:147-156).

    gpurun -- python tests/golden/make_tcnn_loss_curve.py      # writes gpurun_out/tcnn_loss1000_*.npz
    cp gpurun_out/tcnn_loss1000_*.npz tests/golden/

BASELINE north_star: "matching loss curves over 1000 steps".  The fixture holds only the loss curve, the held-out inference
records and tcnn's final predictions on them; the 1000 x batch training records are regenerated from the seed by
`training_data` below (tests/test_gpu_loss_curve.py imports it), so nothing large is committed.
"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
BIN = os.path.join(ROOT, "oracle", "_ref", "tcnn_oracle")
SKY_HALF = np.array([62.317, 42.295, 76.707], dtype=np.float32) / 2      # reference position normalisation

CONFIGS = [  # name, pos, dir, depth, batch, steps, lr
    ("hash_ob_d6", 0, 0, 6, 4096, 1000, 0.01),        # reference default argv (src/main.cu:432-439)
    ("tri_ob_d5", 2, 0, 5, 4096, 1000, 0.01),         # no encoding parameters: isolates the MLP + optimizer
]
N_HELD_OUT = 2048


def radiance_field(rec):
    """smooth RGB in [0, 2] as a function of (position, direction): something the cache can actually learn"""
    x = rec[:, :3].astype(np.float64) - SKY_HALF.astype(np.float64)
    th, ph = rec[:, 3].astype(np.float64), rec[:, 4].astype(np.float64)
    two_pi = 2 * np.pi
    r = 1 + np.sin(two_pi * (x[:, 0] + 2 * x[:, 1])) * np.cos(two_pi * x[:, 2])
    g = 2 * x[:, 0] * x[:, 1] + 0.5 * (1 - x[:, 2]) * (1 + np.cos(two_pi * th)) * 0.5
    b = 1 + 0.8 * np.sin(two_pi * ph) * (x[:, 2] - 0.5) * 2
    return np.clip(np.stack([r, g, b], 1), 0, 2).astype(np.float32)


def records(rng, n):
    rec = rng.random((n, 5), dtype=np.float32)
    rec[:, :3] += SKY_HALF
    return rec


def training_data(seed, batch, steps):
    """(train_in float32[steps*batch][5], train_target float32[steps*batch][3], held_out float32[N_HELD_OUT][5])"""
    rng = np.random.default_rng(seed)
    tin = records(rng, steps * batch)
    held = records(rng, N_HELD_OUT)
    return tin, radiance_field(tin), held


SEED_OFFSETS = (0, 1000, 2000, 3000, 4000, 5000)       # data orders of the multi-seed fixture (offset 0 = the single-curve fixture)


def run_tcnn(name, pos, dr, depth, batch, steps, lr, seed):
    work = f"/tmp/tcnn_loss_{name}_{seed}"
    os.makedirs(work, exist_ok=True)
    tin, tgt, held = training_data(seed, batch, steps)
    held.tofile(work + "/infer_in.f32"); tin.tofile(work + "/train_in.f32"); tgt.tofile(work + "/train_tgt.f32")
    cmd = [BIN, "dump", f"out={work}", f"pos={pos}", f"dir={dr}", f"depth={depth}", f"n_infer={N_HELD_OUT}", f"batch={batch}", f"steps={steps}",
           f"lr={lr}", f"infer_in={work}/infer_in.f32", f"train_in={work}/train_in.f32", f"train_tgt={work}/train_tgt.f32"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        print(res.stdout[-2000:], res.stderr[-2000:], file=sys.stderr)
        raise SystemExit(f"tcnn_oracle failed for {name}")
    losses = np.fromfile(work + "/losses.f32", dtype=np.float32)
    final = np.fromfile(work + "/infer_ema_final.f32", dtype=np.float32).reshape(N_HELD_OUT, 3)
    for f in os.listdir(work):
        os.remove(os.path.join(work, f))
    return losses, final, held


def main():
    """no argument: the single-curve fixtures; `seeds`: tcnn_loss1000_<cfg>_seeds.npz with one curve per data order (the mean over
    the orders is what tests/test_gpu_loss_curve.py holds to the 5 % window bound of SURVEY.md 8d)"""
    out_root = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_root, exist_ok=True)
    args = [a for a in sys.argv[1:]]
    multi = "seeds" in args
    only = set(a for a in args if a != "seeds")
    for name, pos, dr, depth, batch, steps, lr in CONFIGS:
        if only and name not in only:
            continue
        seed = 4242 + pos * 10 + dr
        if multi:
            seeds = [seed + o for o in SEED_OFFSETS]
            curves = np.stack([run_tcnn(name, pos, dr, depth, batch, steps, lr, s)[0] for s in seeds])
            np.savez_compressed(os.path.join(out_root, f"tcnn_loss1000_{name}_seeds.npz"), pos=pos, dir=dr, depth=depth, batch=batch, steps=steps, lr=lr,
                                seeds=np.array(seeds), losses=curves)
            print(json.dumps({"name": name, "seeds": seeds, "loss_first": [float(c[0]) for c in curves]}), file=sys.stderr)
            continue
        losses, final, held = run_tcnn(name, pos, dr, depth, batch, steps, lr, seed)
        np.savez_compressed(os.path.join(out_root, f"tcnn_loss1000_{name}.npz"), pos=pos, dir=dr, depth=depth, batch=batch, steps=steps, lr=lr, seed=seed,
                            losses=losses, held_out=held, infer_ema_final=final, held_out_target=radiance_field(held))
        w = 50
        print(json.dumps({"name": name, "loss_first": float(losses[0]), "loss_window_means": [round(float(losses[i:i + w].mean()), 5) for i in range(0, steps, w * 4)],
                          "final_rel_l2_vs_target": float(np.linalg.norm(final - radiance_field(held)) / np.linalg.norm(radiance_field(held)))}), file=sys.stderr)


if __name__ == "__main__":
    main()
