"""GPU: frame-level parity against the reference's OWN converged frames (reference/<scene>/0.exr, 1920x1080 RGBA32F, produced by
McHpmRenderer with pathLength 64 x 8192 blended frames, reference src/Reference.cpp:443-455, 581-598).  The fixtures are 8x8 block
means of those frames (tests/golden/exr_block8.npz, 240x135; exr_stats.json), generated in the build container from
/root/reference/reference/*/0.exr.  The CUDA path tracer (hpm_mc_render, same tracking code as the NRC passes) renders the same
camera at 240x135 and must reproduce foreground mask, mean radiance and the image itself to Monte-Carlo noise.

Pinned scenes: 0 (directional light) and 4 (directional light + environment light; the EXR was rendered with a uniform white
environment map).  The EXRs of scenes 1, 2 (point light) and 5 were rendered with parameters that differ from the
committed presets (src/AppConfig.cpp:93-150): with the committed values the path tracer -- whose output matches scenes 0 and 4
to 1 % -- is 3.17x / 1.87x / 1.69x brighter and, for scene 5, has a different opacity (alpha 0.989 vs 0.966), see
scripts/diag_frames.py; those three cannot serve as goldens."""
import json
import os

import numpy as np
import pytest

from conftest import ROOT, golden

pytestmark = pytest.mark.gpu

W, H, FRAMES = 240, 135, 192


def quarter_cloud():
    from nrc_hpm_renderer_b200 import volume
    p = os.path.join(ROOT, "data", "wdas_cloud_quarter_u8.npz")
    if not os.path.exists(p):
        pytest.skip("data/wdas_cloud_quarter_u8.npz missing")
    return volume.load_volume(p).data


@pytest.mark.parametrize("scene_id,env", [(0, (0, 0, 0)), (4, (1, 1, 1))])
def test_mc_frames_match_reference_exr(scene_id, env):
    from nrc_hpm_renderer_b200 import Camera, HpmSceneConfig
    from nrc_hpm_renderer_b200.renderer import HpmScene, McHpmRenderer
    stats = json.load(open(os.path.join(ROOT, "tests", "golden", "exr_stats.json")))[str(scene_id)]
    ref = golden("exr_block8.npz")[f"s{scene_id}"].astype(np.float32)          # [135][240][2]: radiance, alpha
    scene = HpmScene(quarter_cloud(), HpmSceneConfig.preset(scene_id), env_color=env)   # Q11: scenes 4/5 were rendered with a white env map
    r = McHpmRenderer(W, H, 64, True, Camera(aspect=1920 / 1080), scene)
    rng = np.random.default_rng(1337)
    for _ in range(FRAMES):
        r.Render(rng.random(4).astype(np.float32))
    img = r.GetImage()
    assert np.isfinite(img).all()
    rad, alpha = img[..., 0], img[..., 3]
    assert np.allclose(img[..., 0], img[..., 1]) and np.allclose(img[..., 0], img[..., 2])      # grey scenes
    # foreground mask (alpha = fraction of frames that scattered)
    fg_ref, fg = ref[..., 1] > 0.02, alpha > 0.02
    assert (fg_ref == fg).mean() >= 0.985
    assert abs(fg.mean() - stats["fg_frac"]) <= 0.02
    both = fg_ref & fg & (ref[..., 1] > 0.5)
    # mean radiance over the medium: the reference's Reference::Result ownMean / refMean (relBias), thesis rBias for MC ~ +-0.01
    rel_bias = (rad[both].mean() - ref[..., 0][both].mean()) / ref[..., 0][both].mean()
    assert abs(rel_bias) <= 0.03, rel_bias
    assert abs(alpha[both].mean() - ref[..., 1][both].mean()) <= 0.02
    # per-pixel relative RMSE: FRAMES blended samples per pixel vs the converged block means
    rmse = np.sqrt(np.mean((rad[both] - ref[..., 0][both]) ** 2)) / ref[..., 0][both].mean()
    assert rmse <= 0.35, rmse
    # background
    if stats["bg_mean"] > 0:
        assert abs(rad[~fg_ref & ~fg].mean() - stats["bg_mean"]) <= 1e-3 * max(1.0, stats["bg_mean"])
    else:
        assert rad[~fg_ref & ~fg].mean() <= 1e-3          # silhouette blocks with alpha < 0.02 carry a little light


def test_nrc_frames_converge_towards_reference():
    """NrcHpmRenderer with online training: the blended NRC image approaches the converged reference (thesis 5.3.3: NRC rBias
    about -0.05 .. -0.08 against the path-traced reference)."""
    from nrc_hpm_renderer_b200 import AppConfig, Camera, HpmSceneConfig
    from nrc_hpm_renderer_b200.nrc import NeuralRadianceCache
    from nrc_hpm_renderer_b200.renderer import HpmScene, NrcHpmRenderer
    ref = golden("exr_block8.npz")["s0"].astype(np.float32)
    app = AppConfig.default()
    app.scene = HpmSceneConfig.preset(0)
    app.log2_train_batch_size, app.train_batch_count = 12, 2
    nrc = NeuralRadianceCache(app)
    scene = HpmScene(quarter_cloud(), app.scene)
    r = NrcHpmRenderer(W, H, False, Camera(aspect=1920 / 1080), app, scene, nrc, parity_q2=False)     # 32-vertex training targets
    rng = np.random.default_rng(7)
    losses = []
    for f in range(300):                                # warm the cache up (online training every frame)
        r.Render(True, rng.random(4).astype(np.float32))
        if f % 50 == 49:
            losses.append(nrc.GetLoss())
    assert np.isfinite(losses).all() and losses[-1] < losses[0] * 1.5
    r.SetBlend(True)
    for f in range(128):
        r.Render(True, rng.random(4).astype(np.float32))
    img = r.GetImage()
    assert np.isfinite(img[..., :3]).all()
    both = (ref[..., 1] > 0.5)
    rel_bias = (img[..., 0][both].mean() - ref[..., 0][both].mean()) / ref[..., 0][both].mean()
    print(f"NRC frames after 428 frames of online training at 240x135: rBias {rel_bias:+.4f} against reference/0/0.exr")
    # thesis 5.3.3 reports -0.05 .. -0.08 for the NRC against the path-traced reference (the cache under-estimates slightly); measured here
    # over repeated runs (the hash-grid gradient is scattered with fp16 atomics, so runs differ): -0.025 .. -0.029.  The band covers
    # the thesis' figures and ours and nothing with the wrong sign or a double-digit loss of energy (round 1 accepted -0.25 .. +0.10).
    assert -0.10 <= rel_bias <= 0.02, rel_bias
