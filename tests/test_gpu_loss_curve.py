"""BASELINE north_star: "matching loss curves over 1000 steps".  The CUDA path trains for 1000 steps on the same records, in the
same order, with the same initial weights (bit-exact, tests/test_gpu_nrc.py) as the reference's own tiny-cuda-nn did when the
fixture tests/golden/tcnn_loss1000_*.npz was recorded on a B200 (generator: tests/golden/make_tcnn_loss_curve.py).
Single steps are compared only through 50-step window means (the trajectories see different rounding: fp32 vs fp16 accumulation,
atomic order).

Tolerance.  SURVEY.md 8(d) proposed 5 % per window; measured on this curve that is tighter than the distance between two FAITHFUL
restatements of the reference: the CPU oracle run for the same 1000 steps in tcnn's fp16-accumulation mode and in fp32 mode
(tests/golden/oracle_loss1000_tri_ob_d5.npz, 275 s each) sits at max |ln(loss/loss_tcnn)| = 0.18 / 0.30 per window (mean 0.075 /
0.10), and the two oracle modes differ from each other by 0.28 -- the loss falls 180x, and during the descent a lag of a few steps
is a two-digit percentage.  The CUDA path (fp32 accumulation) measured 0.36 max / 0.15 mean against tcnn and 0.06 mean against
the fp32 oracle on tri_ob_d5, and 0.13 max / 0.012 mean on hash_ob_d6.  Stated bounds: first loss equal to 1e-3; the first two
windows (before the trajectories decorrelate) within 5 %; every window within |ln| <= 0.45; mean |ln| <= 0.20; same plateau.
After the last step the two caches must predict the held-out records alike."""
import importlib.util
import os

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _generator():
    spec = importlib.util.spec_from_file_location("make_tcnn_loss_curve", os.path.join(ROOT, "tests", "golden", "make_tcnn_loss_curve.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("name", ["hash_ob_d6", "tri_ob_d5"])
def test_loss_curve_1000_steps_vs_tcnn(name):
    import torch
    from nrc_hpm_renderer_b200 import AppConfig
    from nrc_hpm_renderer_b200.nrc import NeuralRadianceCache
    path = os.path.join(ROOT, "tests", "golden", f"tcnn_loss1000_{name}.npz")
    if not os.path.exists(path):
        pytest.skip("fixture not recorded yet")
    z = np.load(path)
    B, steps = int(z["batch"]), int(z["steps"])
    tin, tgt, held = _generator().training_data(int(z["seed"]), B, steps)
    assert np.array_equal(held, z["held_out"])                     # same generator, same records
    app = AppConfig.default()
    app.pos_enc_id, app.dir_enc_id, app.nn_depth, app.learning_rate = int(z["pos"]), int(z["dir"]), int(z["depth"]), float(z["lr"])
    c = NeuralRadianceCache(app)
    d_in, d_tgt = torch.from_numpy(tin).cuda(), torch.from_numpy(tgt).cuda()
    losses = np.empty(steps, np.float32)
    for s in range(steps):
        c.training_step(d_in[s * B:(s + 1) * B], d_tgt[s * B:(s + 1) * B], B, True)
        losses[s] = c.GetLoss()
    ref = z["losses"]
    assert np.all(np.isfinite(losses))
    assert abs(losses[0] - ref[0]) <= 1e-3 * ref[0]                 # identical weights and records: the first loss is the same number
    w = 50
    ln = np.array([np.log(losses[i:i + w].mean() / ref[i:i + w].mean()) for i in range(0, steps, w)])
    assert np.abs(ln[:2]).max() <= 0.05, ln[:2]
    assert np.abs(ln).max() <= 0.45 and np.abs(ln).mean() <= 0.20, (np.abs(ln).max(), np.abs(ln).mean())
    assert ref[-w:].mean() < 0.25 * ref[:w].mean() and losses[-w:].mean() < 0.25 * losses[:w].mean()      # both actually learned
    out = torch.zeros((len(held), 3), dtype=torch.float32, device="cuda")
    c.inference(torch.from_numpy(held).cuda(), out, len(held), use_ema=True)
    torch.cuda.synchronize()
    ours, theirs, truth = out.cpu().numpy(), z["infer_ema_final"], z["held_out_target"]
    e_ours = np.linalg.norm(ours - truth) / np.linalg.norm(truth)
    e_theirs = np.linalg.norm(theirs - truth) / np.linalg.norm(truth)
    assert e_ours <= 1.15 * e_theirs + 0.01, (e_ours, e_theirs)   # as good a fit as the reference's
    assert np.linalg.norm(ours - theirs) / np.linalg.norm(theirs) <= 2.0 * max(e_ours, e_theirs) + 0.01


@pytest.mark.parametrize("name,max_ln,mean_ln", [("hash_ob_d6", 0.05, 0.01), ("tri_ob_d5", 0.10, 0.03)])
def test_loss_curve_mean_over_data_orders_vs_tcnn(name, max_ln, mean_ln):
    """loss curves within 5 % per 50-step window.  A single trajectory cannot be held to that (docstring above: two
    faithful restatements differ by more), but the EXPECTED curve can: tests/golden/tcnn_loss1000_<cfg>_seeds.npz holds the
    reference's own tiny-cuda-nn curves for six data orders (make_tcnn_loss_curve.py seeds, recorded on a B200); the CUDA path
    trains on the same six orders and the window means, averaged over the orders, are compared.  Measured (round 2, B200):
    hash_ob_d6 (the reference default) max |ln| 0.030, mean 0.003 -> bound 5 %, the contract; tri_ob_d5 (no encoding parameters,
    the loss falls 180x) max 0.074, mean 0.018, while tcnn's own order-to-order spread of a window mean is sigma(ln) = 0.3 there,
    i.e. a standard error of 0.12 for six orders -> bound 10 %."""
    import torch
    from nrc_hpm_renderer_b200 import AppConfig
    from nrc_hpm_renderer_b200.nrc import NeuralRadianceCache
    path = os.path.join(ROOT, "tests", "golden", f"tcnn_loss1000_{name}_seeds.npz")
    if not os.path.exists(path):
        pytest.skip("fixture not recorded yet")
    z = np.load(path)
    B, steps = int(z["batch"]), int(z["steps"])
    ours = []
    for seed in z["seeds"]:
        tin, tgt, _ = _generator().training_data(int(seed), B, steps)
        app = AppConfig.default()
        app.pos_enc_id, app.dir_enc_id, app.nn_depth, app.learning_rate = int(z["pos"]), int(z["dir"]), int(z["depth"]), float(z["lr"])
        c = NeuralRadianceCache(app)
        d_in, d_tgt = torch.from_numpy(tin).cuda(), torch.from_numpy(tgt).cuda()
        losses = np.empty(steps, np.float32)
        for s in range(steps):
            c.training_step(d_in[s * B:(s + 1) * B], d_tgt[s * B:(s + 1) * B], B, True)
            losses[s] = c.GetLoss()
        ours.append(losses)
        c.Destroy()
    ours, ref = np.stack(ours), z["losses"]
    assert np.all(np.isfinite(ours))
    assert np.max(np.abs(ours[:, 0] - ref[:, 0]) / ref[:, 0]) <= 1e-3          # identical weights and records: same first loss, every order
    w = 50
    om = ours.reshape(len(ours), -1, w).mean(2).mean(0); tm = ref.reshape(len(ref), -1, w).mean(2).mean(0)
    ln = np.log(om / tm)
    print(f"{name}: mean over {len(ours)} data orders, per-window ln(loss/loss_tcnn): max {np.abs(ln).max():.3f} mean {np.abs(ln).mean():.3f}")
    assert np.abs(ln).max() <= max_ln and np.abs(ln).mean() <= mean_ln, (np.abs(ln).max(), np.abs(ln).mean())
