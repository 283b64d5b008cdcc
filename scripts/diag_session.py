"""Development tool (GPU): diagnostics behind two parity tests -- compact vs all-records frames, and the 1000-step loss curves."""
import importlib.util, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import oracle as O
from nrc_hpm_renderer_b200 import AppConfig
from nrc_hpm_renderer_b200.nrc import NeuralRadianceCache

def frames():
    import test_gpu_tracker as T
    from nrc_hpm_renderer_b200 import renderer as R
    O.build()
    W, H = 128, 64
    imgs = []
    for compact in (True, False, True):
        app = AppConfig.default(); app.log2_train_batch_size, app.train_batch_count, app.log2_infer_batch_size = 9, 2, 12
        nrc = NeuralRadianceCache(app)
        r, *_ = T.setup(0, W, H, O, train_pixels=1024, compact=compact, nrc=nrc)
        fr = []
        for f in range(3):
            r.Render(True, T.FR + np.float32(0.05 * f))
            fr.append(r.GetImage().copy())
        imgs.append(fr)
    for name, (a, b) in {"compact_vs_all": (imgs[0], imgs[1]), "compact_vs_compact": (imgs[0], imgs[2])}.items():
        for f in range(3):
            d = np.abs(a[f] - b[f]); d = d[np.isfinite(d)]
            print(json.dumps({"cmp": name, "frame": f, "n_diff": int((d != 0).sum()), "max_abs": float(d.max()) if d.size else 0.0, "mean_img": float(np.nanmean(a[f][..., :3]))}))

def losses(name):
    spec = importlib.util.spec_from_file_location("g", os.path.join(ROOT, "tests", "golden", "make_tcnn_loss_curve.py"))
    g = importlib.util.module_from_spec(spec); spec.loader.exec_module(g)
    z = np.load(os.path.join(ROOT, "tests", "golden", f"tcnn_loss1000_{name}.npz"))
    B, steps = int(z["batch"]), int(z["steps"])
    tin, tgt, held = g.training_data(int(z["seed"]), B, steps)
    app = AppConfig.default()
    app.pos_enc_id, app.dir_enc_id, app.nn_depth, app.learning_rate = int(z["pos"]), int(z["dir"]), int(z["depth"]), float(z["lr"])
    c = NeuralRadianceCache(app)
    d_in, d_tgt = torch.from_numpy(tin).cuda(), torch.from_numpy(tgt).cuda()
    ours = np.empty(steps, np.float32)
    for s in range(steps):
        c.training_step(d_in[s * B:(s + 1) * B], d_tgt[s * B:(s + 1) * B], B, True)
        ours[s] = c.GetLoss()
    ref = z["losses"]
    w = 50
    print(json.dumps({"name": name, "first": [float(ours[0]), float(ref[0])],
                      "windows": [[round(float(ours[i:i + w].mean()), 5), round(float(ref[i:i + w].mean()), 5)] for i in range(0, steps, w)]}))
    np.savez(os.path.join(ROOT, "gpurun_out", f"ours_loss_{name}.npz"), ours=ours, ref=ref)

if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    frames()
    for n in ("hash_ob_d6", "tri_ob_d5"):
        losses(n)
