"""GPU: the path-regeneration schedule of the gen_rays pass (csrc/hpm_wavefront.cuh: primary rays in one coherent launch, then
persistent warps in which a lane whose path has ended takes the next path from a queue; early launches hand their stragglers to the
next one) against the pixel-per-thread kernel, which the other tracker tests hold to the CPU oracle.  Same per-pixel arithmetic, same
per-pixel RNG stream: every buffer must be IDENTICAL bit for bit, the density lookup counter included; only the order of the compacted
record list may differ (it is appended with atomics in both modes).

Cases follow the reference's scene presets (src/AppConfig.cpp:93-150: which lights are on) and the path-length parameters of
gen_rays.comp:39-42 (fixed length, Russian roulette, longer than any queue round)."""
import os

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

FR = np.array([0.3183, 0.7071, 0.1234, 0.9876], np.float32)


def quarter_cloud():
    from nrc_hpm_renderer_b200 import volume
    p = os.path.join(ROOT, "data", "wdas_cloud_quarter_u8.npz")
    if not os.path.exists(p):
        pytest.skip("data/wdas_cloud_quarter_u8.npz missing")
    return volume.load_volume(p).data


def run_mode(r, mode, fr):
    from nrc_hpm_renderer_b200 import renderer as R
    r.set_tracker_mode(mode)
    r.pass_gen_rays(fr); r.sync()
    out = {k: r.read(b).copy() for k, b in (("info", R.BUF_PRIMARY_INFO), ("color", R.BUF_PRIMARY_COLOR), ("origin", R.BUF_NRC_ORIGIN), ("dir", R.BUF_NRC_DIR),
                                             ("infer_in", R.BUF_INFER_INPUT), ("filter", R.BUF_INFER_FILTER), ("counters", R.BUF_COUNTERS))}
    return out


def assert_identical(a, b):
    for k in ("info", "color", "origin", "dir", "infer_in"):
        assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), k          # bit patterns: NaN phi (Q5) must match too
    assert np.array_equal(a["filter"], b["filter"])
    assert int(a["counters"][0]) == int(b["counters"][0]), (a["counters"], b["counters"])     # density lookups
    assert int(a["counters"][2]) == int(b["counters"][2]) == int((a["info"] == 1.0).sum())    # compacted records


CASES = [
    # scene preset, overrides, primaryRayLength, primaryRayProb, (W, H), column strip
    (0, {}, 1, 0.0, (480, 270), None),                                     # default argv: directional light only
    (4, {}, 1, 0.0, (480, 270), None),                                     # directional + environment
    (1, {}, 2, 0.0, (320, 200), None),                                     # point light only, three bounces
    (4, {"point_light_strength": 32.0}, 0, 0.0, (333, 177), (40, 301)),    # all three lights, ragged size, a column strip of a tile-partitioned frame
    (5, {}, 4, 0.75, (480, 270), None),                                    # BASELINE config 4: dense medium, Russian roulette
    (0, {"dir_light_strength": 0.0}, 1, 0.5, (256, 256), None),            # no light at all
    (3, {}, 30, 0.0, (160, 90), None),                                     # long fixed paths: several spill rounds
]


@pytest.mark.parametrize("preset,over,length,prob,size,strip", CASES)
def test_regeneration_is_bit_identical_to_pixel_per_thread(preset, over, length, prob, size, strip):
    from nrc_hpm_renderer_b200 import AppConfig, Camera, HpmSceneConfig
    from nrc_hpm_renderer_b200 import renderer as R
    W, H = size
    grid = quarter_cloud()
    app = AppConfig.default()
    app.scene = HpmSceneConfig.preset(preset)
    for k, v in over.items():
        assert hasattr(app.scene, k), k
        setattr(app.scene, k, v)
    app.primary_ray_length, app.primary_ray_prob = length, prob
    scene = R.HpmScene(grid, app.scene)
    cam = Camera(aspect=W / H)
    kw = {} if strip is None else {"x_begin": strip[0], "x_end": strip[1]}
    cfg = R.make_render_config(W, H, app, train_pixels=0, **kw)
    r = R.NrcHpmRenderer(W, H, False, cam, app, scene, None, render_config=cfg)
    for fr in (FR, FR[::-1].copy()):
        a = run_mode(r, 1, fr)
        b = run_mode(r, 2, fr)
        assert a["info"].sum() > 0.05 * (W * H if strip is None else (strip[1] - strip[0]) * H)       # the case exercises the volume
        assert_identical(a, b)
        c = run_mode(r, 2, fr)                                                                       # and reproducible
        assert_identical(b, c)


def test_regeneration_frames_equal_pixel_per_thread_frames():
    """whole frame (tracking, compacted inference, compositing) in both modes: identical output images -- the compacted record list
    differs in order only, and a record's inference result does not depend on its place in the list.  (No training between the frames:
    the hash-grid gradient is scattered with fp16 atomics, whose order is not reproducible from run to run in EITHER mode.)"""
    from nrc_hpm_renderer_b200 import AppConfig, Camera, HpmSceneConfig
    from nrc_hpm_renderer_b200 import nrc as N, renderer as R
    W, H = 480, 270
    grid = quarter_cloud()
    imgs = []
    for mode in (1, 2):
        app = AppConfig.default()
        app.scene = HpmSceneConfig.preset(0)
        cache = N.NeuralRadianceCache(app)
        cache.set_ema(cache.get_params(N.MASTER))              # Inference() reads the EMA weights, zero before the first step (Q7)
        scene = R.HpmScene(grid, app.scene)
        r = R.NrcHpmRenderer(W, H, False, Camera(aspect=W / H), app, scene, cache)
        r.set_tracker_mode(mode)
        for f in range(2):
            r.Render(False, FR * (1 + f)); r.sync()
        imgs.append(r.GetImage().copy())
    assert np.abs(imgs[0][..., :3]).max() > 0
    assert np.array_equal(imgs[0].view(np.uint32), imgs[1].view(np.uint32))
