"""GPU parity of the tracking passes (through the C ABI) against the CPU oracle (oracle/hpm_oracle.cpp) on the same scene,
camera and frame seed.  The RNG is bit-exact; + - * / sqrt are IEEE-exact on both sides (-fmad=false / -ffp-contract=off),
so a pixel's path only departs where CUDA's logf/sinf/cosf/acosf/powf/atan2f round differently from glibc's (<= 2 ulp):
the bar is an identical scatter decision for >= 99 % of the pixels and matching values on those pixels.

NaN pixels are legitimate reference behaviour: random.glsl's state has an absorbing value (hash(0) == 0, so a state whose
23 mantissa bits come out zero stays 0.0 forever, probability 2^-23 per draw); such a pixel draws u == 0 for the rest of
its path and NewRayDir evaluates acos(-1.0000004) (dir_gen.glsl:44-47).  Oracle and CUDA path agree on those pixels, so
the comparisons below are NaN-aware."""
import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu

FR = np.array([0.3183, 0.7071, 0.1234, 0.9876], np.float32)


def setup(scene_id, W, H, oracle, *, train_pixels=1024, ring_frac=1.0, prl=1, prob=0.0, spp=1, trl=1, env=(0, 0, 0), compact=True, nrc=None, blend=False,
          log2_infer_batch=None, grid=None, density=None):
    from nrc_hpm_renderer_b200 import AppConfig, Camera, HpmSceneConfig
    from nrc_hpm_renderer_b200.renderer import HpmScene, NrcHpmRenderer, make_render_config
    if grid is None:
        grid = np.ascontiguousarray(golden("wdas_cloud_sixteenth_u8.npz")["data"])
    app = AppConfig.default()
    app.scene = HpmSceneConfig.preset(scene_id)
    app.primary_ray_length, app.primary_ray_prob, app.train_spp, app.train_ring_buf_size = prl, prob, spp, ring_frac
    app.train_ray_length = trl
    if log2_infer_batch is not None:
        app.log2_infer_batch_size = log2_infer_batch          # the renderer's filter granularity == the cache's batch size (one AppConfig)
    cam = Camera(aspect=W / H)
    scene = HpmScene(grid, app.scene, env_color=env, density=density)
    cfg = make_render_config(W, H, app, blend=blend, compact_inference=compact, train_pixels=train_pixels, parity_q2=False)
    r = NrcHpmRenderer(W, H, blend, cam, app, scene, nrc, render_config=cfg)
    d = scene.desc
    osc = oracle.make_scene(grid, tuple(d.sky_size), d.density_factor, d.g, tuple(d.dir_light_dir), d.dir_light_strength, tuple(d.point_pos),
                            d.point_strength, tuple(d.point_color), d.env_strength, tuple(d.env_color))
    ocfg = oracle.make_config(W, H, cfg.train_width, cfg.train_height, cfg.train_x_dist, cfg.train_y_dist, cfg.train_spp, cfg.primary_ray_length,
                              cfg.primary_ray_prob, cfg.train_ring_size, cfg.train_ray_length, cfg.infer_batch_size)
    ocam = oracle.make_camera(cam.inv_proj_view, cam.pos)
    return r, osc, ocfg, ocam


# last case = BASELINE config 4: the dense medium (scene-5 density 1.6, src/AppConfig.cpp:138-145) with the thesis' worst-case termination
# (primaryRayLength 4, primaryRayProb .75: gen_rays.comp:39-42) -- long paths, the divergence / compaction stress case
@pytest.mark.parametrize("scene_id,prl,prob,env", [(0, 1, 0.0, (0, 0, 0)), (1, 1, 0.0, (0, 0, 0)), (4, 2, 0.5, (1, 1, 1)), (2, 0, 0.0, (0, 0, 0)), (5, 4, 0.75, (1, 1, 1))])
def test_gen_rays_vs_oracle(scene_id, prl, prob, env, oracle_lib):
    from nrc_hpm_renderer_b200 import renderer as R
    W, H = 160, 96
    r, osc, ocfg, ocam = setup(scene_id, W, H, oracle_lib, prl=prl, prob=prob, env=env)
    r.pass_gen_rays(FR)
    ref = oracle_lib.gen_rays(osc, ocfg, ocam, FR)
    info = r.read(R.BUF_PRIMARY_INFO)
    color = r.read(R.BUF_PRIMARY_COLOR).reshape(-1, 4)
    same = info == ref["info"]
    assert same.mean() >= 0.99, same.mean()
    assert 0.05 < info.mean() < 0.9
    # pixels whose whole path agreed: colour / throughput / vertex agree to float noise
    cd = np.abs(color - ref["color"]).max(axis=1)
    tol = 1e-3 * np.maximum(1.0, np.abs(ref["color"]).max(axis=1))
    agree = same & (cd <= tol)
    assert agree.mean() >= 0.97, agree.mean()
    org = r.read(R.BUF_NRC_ORIGIN).reshape(-1, 3); dr = r.read(R.BUF_NRC_DIR).reshape(-1, 3)
    sc = agree & (info == 1.0)
    assert np.nanpercentile(np.abs(org[sc] - ref["origin"][sc]).max(axis=1), 98) <= 1e-3
    assert np.nanpercentile(np.abs(dr[sc] - ref["dir"][sc]).max(axis=1), 98) <= 1e-3
    # fused prep_infer_rays: records at x*H+y, zeros elsewhere, filter flags, compacted index list
    rec, filt = oracle_lib.prep_infer(osc, ocfg, dict(info=info, origin=org, dir=dr))
    got = r.read(R.BUF_INFER_INPUT).reshape(-1, 5)
    both = np.isfinite(rec)
    assert np.array_equal(np.isfinite(got), both)
    assert np.max(np.abs(got[both] - rec[both])) <= 2e-6 * 40      # atan2/acos rounding on ~[0,40] values
    assert np.array_equal(r.read(R.BUF_INFER_FILTER), filt)
    cnt = r.read(R.BUF_COUNTERS)
    lin = (np.arange(W)[None, :] * H + np.arange(H)[:, None]).reshape(-1)       # pixel (x,y) at y*W+x -> record x*H+y
    assert int(cnt[2]) == int((info == 1.0).sum())
    # density fetch count (roofline accounting) within 1 % of the oracle's
    assert abs(int(cnt[0]) - ref["lookups"]) / ref["lookups"] <= 0.01


def test_gen_rays_statistics_match_oracle(oracle_lib):
    """image-level agreement over several seeds: mean radiance and scatter fraction (paths that diverged are still unbiased)"""
    from nrc_hpm_renderer_b200 import renderer as R
    W, H = 96, 64
    r, osc, ocfg, ocam = setup(0, W, H, oracle_lib)
    g, o = [], []
    for s in range(4):
        fr = np.array([0.1 + 0.2 * s, 0.9 - 0.1 * s, 0.5, 0.25 * s], np.float32)
        r.pass_gen_rays(fr)
        g.append(np.nanmean(r.read(R.BUF_PRIMARY_COLOR).reshape(-1, 4)[:, 0]))
        o.append(np.nanmean(oracle_lib.gen_rays(osc, ocfg, ocam, fr)["color"][:, 0]))
    assert abs(np.mean(g) - np.mean(o)) / np.mean(o) <= 0.02


@pytest.mark.parametrize("ring_frac,trl,spp", [(1.0, 1, 1), (0.5, 3, 2)])
def test_prep_train_vs_oracle(ring_frac, trl, spp, oracle_lib):
    from nrc_hpm_renderer_b200 import renderer as R
    W, H = 128, 96
    r, osc, ocfg, ocam = setup(0, W, H, oracle_lib, train_pixels=1536, ring_frac=ring_frac, trl=trl, spp=spp)
    ring = oracle_lib.new_ring(ocfg)
    assert np.array_equal(r.read(R.BUF_TRAIN_RING), ring)       # CreateNrcTrainRingBuffer initial state
    for frame in range(3):                                      # three frames: the ring fills, wraps and is popped
        fr = FR + np.float32(0.01 * frame)
        r.pass_gen_rays(fr)
        rays = dict(info=r.read(R.BUF_PRIMARY_INFO), origin=r.read(R.BUF_NRC_ORIGIN).reshape(-1, 3), dir=r.read(R.BUF_NRC_DIR).reshape(-1, 3))
        r.write(R.BUF_TRAIN_RING, ring)                         # same ring state on both sides
        r.pass_prep_train(fr)
        tin, tgt, lookups = oracle_lib.prep_train(osc, ocfg, rays, fr, ring)
        got_ring = r.read(R.BUF_TRAIN_RING)
        assert np.array_equal(got_ring, ring)                   # head, tail and every stored ray: bit-exact (pure data movement)
        gin = r.read(R.BUF_TRAIN_INPUT).reshape(-1, 5); gt = r.read(R.BUF_TRAIN_TARGET).reshape(-1, 3)
        both = np.isfinite(tin)
        assert np.array_equal(np.isfinite(gin), both)
        assert np.max(np.abs(gin[both] - tin[both])) <= 1e-4
        close = np.abs(gt - tgt).max(axis=1) <= 1e-3 * np.maximum(1.0, np.abs(tgt).max(axis=1))
        assert close.mean() >= 0.97, close.mean()
        assert np.all(gt <= 8.0)
        cnt = r.read(R.BUF_COUNTERS)
        assert abs(int(cnt[1]) - lookups) / max(lookups, 1) <= 0.02


def test_composite_bit_exact(oracle_lib):
    from nrc_hpm_renderer_b200 import renderer as R
    W, H = 64, 32
    r, osc, ocfg, ocam = setup(0, W, H, oracle_lib, blend=True)
    r.pass_gen_rays(FR)
    rays = dict(color=r.read(R.BUF_PRIMARY_COLOR).reshape(-1, 4), info=r.read(R.BUF_PRIMARY_INFO))
    rng = np.random.default_rng(0)
    nrc_out = (rng.random((W * H, 3), dtype=np.float32) - 0.3).astype(np.float32)
    r.write(R.BUF_INFER_OUTPUT, nrc_out)
    r.SetBlend(False)
    r.pass_composite()
    ref1 = np.zeros((W * H, 4), np.float32)
    oracle_lib.render(ocfg, rays, nrc_out, 1, 1.0, ref1)
    assert np.array_equal(r.read(R.BUF_OUTPUT).reshape(-1, 4), ref1, equal_nan=True)


def test_mc_render_vs_oracle(oracle_lib):
    from nrc_hpm_renderer_b200 import renderer as R
    W, H = 96, 64
    r, osc, ocfg, ocam = setup(1, W, H, oracle_lib, train_pixels=0)
    r.mc_render(FR, 4)
    out = np.zeros((W * H, 4), np.float32)
    oracle_lib.mc_render(osc, ocfg, ocam, FR, 4, 1.0, out)
    got = r.read(R.BUF_OUTPUT).reshape(-1, 4)
    assert (got[:, 3] == out[:, 3]).mean() >= 0.99
    close = np.abs(got - out).max(axis=1) <= 1e-3 * np.maximum(1.0, np.abs(out).max(axis=1))
    assert close.mean() >= 0.95


def test_frame_compact_equals_reference_mode(oracle_lib):
    """Render() with warp-compacted inference == Render() in reference mode (host filter, all records), pixel for pixel, for the
    same cache parameters.  Frame 0 shows the primary estimate only because the EMA weights are still zero (Q7).  Across *training*
    frames the two modes may drift apart in the last bits: the hash-grid gradient is summed with fp16 atomics whose order depends on
    what ran before (same property as tcnn's kernel_grid_backward) and Adam's first steps amplify a sign flip of a ~0 gradient; so
    trained frames are compared statistically and the exact comparison is made with the parameters copied across and train=false."""
    from nrc_hpm_renderer_b200 import AppConfig, renderer as R
    from nrc_hpm_renderer_b200 import nrc as N
    W, H = 128, 64
    imgs, caches, renderers = [], [], []
    for compact in (True, False):
        app = AppConfig.default(); app.log2_train_batch_size, app.train_batch_count, app.log2_infer_batch_size = 9, 2, 12
        nrc = N.NeuralRadianceCache(app)
        r, *_ = setup(0, W, H, oracle_lib, train_pixels=1024, compact=compact, nrc=nrc, log2_infer_batch=12)
        frames = []
        for f in range(3):
            r.Render(True, FR + np.float32(0.05 * f))
            frames.append(r.GetImage().copy())
        imgs.append(frames); caches.append(nrc); renderers.append(r)
        ms = r.EvaluateTimestampQueries()
        assert ms["total"] > 0 and ms["gen_rays"] > 0
        assert np.isfinite(nrc.GetLoss())
        if compact:
            assert np.all((frames[0][..., :3] >= 0) | np.isnan(frames[0][..., :3]))
    assert np.array_equal(imgs[0][0], imgs[1][0], equal_nan=True)                 # frame 0: no cache term yet, bit-identical
    for a, b in zip(imgs[0][1:], imgs[1][1:]):                                    # trained frames: same up to atomic-order drift
        d = np.abs(a - b)[np.isfinite(a) & np.isfinite(b)]
        assert d.mean() <= 2e-3 and d.max() <= 0.2, (d.mean(), d.max())
    # identical parameters, no training: the two inference modes must agree bit for bit
    caches[1].set_params(caches[0].get_params(N.MASTER))
    caches[1].set_ema(caches[0].get_params(N.EMA))
    fr = FR + np.float32(0.3)
    for r in renderers:
        r.Render(False, fr)
    a, b = renderers[0].GetImage(), renderers[1].GetImage()
    assert np.array_equal(a, b, equal_nan=True)
    prim = renderers[0].read(R.BUF_PRIMARY_COLOR).reshape(H, W, 4)
    assert np.any(a[..., :3] != prim[..., :3])                                    # the cache term is really there
    # a renderer whose filter granularity differs from the cache's batch size would silently skip batches: refused
    with pytest.raises(Exception, match="infer_batch_size"):
        setup(0, W, H, oracle_lib, train_pixels=1024, compact=False, nrc=caches[1], log2_infer_batch=13)
    # frame 0: NRC output is exactly zero -> image == primary estimate; later frames add a non-negative cache term
    assert np.any(imgs[0][2] != imgs[0][0])


def test_tracker_logf_is_the_library_logf_on_every_reachable_argument():
    """The Woodcock loops call a branch-free logf (no denormal / zero / inf / NaN paths); it must equal CUDA's logf bit for bit on
    all 2^23 arguments 1 - u the RNG can produce."""
    import ctypes as C
    from nrc_hpm_renderer_b200 import _lib
    n = C.c_uint64(123)
    _lib.check(_lib.lib().hpm_selftest_logf(C.byref(n)))
    assert n.value == 0


def test_pipelined_training_keeps_the_order_of_effects(oracle_lib):
    """pipeline_train: Train() of frame N runs on its own stream underneath the tracking of frame N+1.  With an encoding without
    parameters (no atomics: training is deterministic) the frames and the final parameters are bit-identical to the serial renderer."""
    from nrc_hpm_renderer_b200 import AppConfig, Camera, HpmSceneConfig
    from nrc_hpm_renderer_b200 import nrc as N, renderer as R
    grid = np.ascontiguousarray(golden("wdas_cloud_sixteenth_u8.npz")["data"])
    W, H = 128, 64
    runs = []
    for pipelined in (False, True):
        app = AppConfig.default(); app.scene = HpmSceneConfig.preset(0)
        app.pos_enc_id, app.log2_train_batch_size, app.train_batch_count = 2, 9, 2
        nrc = N.NeuralRadianceCache(app)
        scene = R.HpmScene(grid, app.scene)
        cfg = R.make_render_config(W, H, app, train_pixels=1024, pipeline_train=pipelined, parity_q2=False)
        r = R.NrcHpmRenderer(W, H, False, Camera(aspect=W / H), app, scene, nrc, render_config=cfg)
        frames = []
        for f in range(6):
            r.Render(True, FR + np.float32(0.07 * f))
            frames.append(r.GetImage().copy())
        r.sync()
        runs.append((frames, nrc.get_params(N.MASTER), nrc.get_params(N.EMA), nrc.GetLoss()))
    for a, b in zip(runs[0][0], runs[1][0]):
        assert np.array_equal(a, b, equal_nan=True)
    assert np.array_equal(runs[0][1], runs[1][1]) and np.array_equal(runs[0][2], runs[1][2]) and runs[0][3] == runs[1][3]
    assert np.any(runs[0][0][5] != runs[0][0][0])
