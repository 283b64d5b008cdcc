"""CPU: host-side mirrors of the reference interface (AppConfig, CalcTrainSubset, camera matrices, VDB reader, strips)."""
import math
import os

import numpy as np
import pytest

from nrc_hpm_renderer_b200 import AppConfig, Camera, HpmSceneConfig, calc_train_subset, encoding_json, sky_size
from nrc_hpm_renderer_b200 import volume
from nrc_hpm_renderer_b200.config import DEFAULT_ARGV
from nrc_hpm_renderer_b200.parallel import column_strips

from conftest import ROOT, golden


def test_appconfig_default_argv():
    a = AppConfig.from_argv(DEFAULT_ARGV)                     # reference src/main.cu:432-439
    assert (a.loss_fn, a.optimizer, a.learning_rate, a.ema_decay) == ("RelativeL2Luminance", "Adam", 0.01, 0.99)
    assert (a.pos_enc_id, a.dir_enc_id, a.nn_width, a.nn_depth) == (0, 0, 64, 6)
    assert (a.infer_batch_size, a.train_batch_size, a.train_batch_count) == (1 << 21, 1 << 14, 4)
    assert a.scene.id == 4 and a.scene.dir_light_strength == 8.0 and a.scene.hdr_env_map_strength == 0.1 and a.scene.density == 0.6
    assert (a.train_ring_buf_size, a.train_spp, a.primary_ray_length, a.primary_ray_prob, a.train_ray_length) == (1.0, 1, 1, 0.0, 32)
    j = a.model_json()
    assert j["optimizer"] == {"otype": "EMA", "decay": 0.99, "nested": {"otype": "Adam", "learning_rate": 0.01}}
    assert j["encoding"]["nested"][0]["otype"] == "HashGrid" and j["encoding"]["nested"][1] == {"otype": "OneBlob", "n_dims_to_encode": 2, "n_bins": 4}
    assert j["network"] == {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None", "n_neurons": 64, "n_hidden_layers": 6}
    assert a.get_name().startswith("RelativeL2Luminance_Adam_0.010000_0.990000_0_0_64_6_21_14_4_4_1.000000_1_1_0.000000_32")


def test_appconfig_errors():
    with pytest.raises(RuntimeError):
        AppConfig.from_argv(DEFAULT_ARGV[:-1])                # exactly 18 entries (src/AppConfig.cpp:156)
    with pytest.raises(RuntimeError):
        encoding_json(4, 0)
    with pytest.raises(RuntimeError):
        encoding_json(0, 3)
    with pytest.raises(RuntimeError):
        HpmSceneConfig.preset(6)


def test_calc_train_subset():
    t = calc_train_subset(1920, 1080, 4 << 14)               # 256x256, XDist 7, YDist 4
    assert (t.train_width, t.train_height, t.x_dist, t.y_dist) == (256, 256, 7, 4)
    t = calc_train_subset(256, 256, 1 << 14)
    assert (t.train_width, t.train_height, t.x_dist, t.y_dist) == (128, 128, 2, 2)
    t = calc_train_subset(1920, 1080, 6 << 14)               # 98304 = 384 x 256, landscape gets the wider side
    assert t.train_width * t.train_height == 98304 and t.train_width >= t.train_height
    with pytest.raises(RuntimeError):
        calc_train_subset(64, 64, 7919)                      # prime: no factorisation (reference logs an error)


def test_sky_size_and_camera():
    s = sky_size((498, 338, 613))
    assert np.allclose(s, (62.317, 42.295, 76.707), atol=2e-3)
    cam = Camera(aspect=1920 / 1080)
    pv, inv = cam.matrices()
    PV, INV = pv.reshape(4, 4).T.astype(np.float64), inv.reshape(4, 4).T.astype(np.float64)
    assert np.allclose(PV @ INV, np.eye(4), atol=1e-4)
    # centre pixel looks down -x from (64,0,0): uv = 0.5 -> ndc 0 (gen_rays.comp:60-72)
    p = INV @ np.array([0.0, 0.0, 0.0, 1.0]); p = p[:3] / p[3]
    d = p - np.array(cam.pos); d /= np.linalg.norm(d)
    assert np.allclose(d, (-1, 0, 0), atol=1e-5)
    # glm conventions: +y up, no Vulkan flip: ndc y = +1 maps to world +y
    p = INV @ np.array([0.0, 1.0, 0.0, 1.0]); p = p[:3] / p[3]
    assert p[1] > 0
    d = p - np.array(cam.pos); d /= np.linalg.norm(d)
    assert math.isclose(math.degrees(math.acos(-d[0])), 30.0, abs_tol=0.01)     # top edge of a 60 degree vertical fov


def test_sixteenth_cloud_fixture():
    z = golden("wdas_cloud_sixteenth_u8.npz")
    g = z["data"]
    assert g.dtype == np.uint8 and g.shape == (154, 86, 126) and g.max() == 255       # max value exactly 1.0 (Texture3D.cpp:74)
    assert int(z["active_voxels"]) == 418490 or int(z["active_voxels"]) > 400000
    ref = "/root/reference/data/volume/wdas_cloud_sixteenth.vdb"
    if os.path.exists(ref):                                                               # only in the build container
        v = volume.read_vdb_dense(ref)
        assert v.dims == (126, 86, 154) and np.array_equal(v.data, g) and v.max_value == 1.0


def test_vdb_reader_rejects_garbage(tmp_path):
    p = tmp_path / "x.vdb"
    p.write_bytes(b"not a vdb file at all" * 10)
    with pytest.raises(Exception):
        volume.read_vdb_dense(str(p))


def test_synthetic_cloud_and_roundtrip(tmp_path):
    v = volume.synthetic_cloud((32, 24, 40), seed=3)
    assert v.dims == (32, 24, 40) and v.data.max() == 255
    volume.save_volume(v, str(tmp_path / "v.npz"))
    w = volume.load_volume(str(tmp_path / "v.npz"))
    assert np.array_equal(v.data, w.data) and w.dims == v.dims


def test_column_strips():
    s = column_strips(1920, 8)
    assert s[0][0] == 0 and s[-1][1] == 1920 and all(a[1] == b[0] for a, b in zip(s, s[1:]))
    assert all((e - b) % 64 == 0 for b, e in s[:-1])
    assert column_strips(1920, 1) == [(0, 1920)]
    s = column_strips(3840, 4)
    assert [e - b for b, e in s] == [960] * 4


def test_tile_render_configs_partition_frame_and_train_lattice():
    """screen tiles by column strips that follow the frame's train-pixel lattice; every rank trains on 1/world of
    each batch.  Pure host logic (ctypes struct only)."""
    from nrc_hpm_renderer_b200 import AppConfig
    from nrc_hpm_renderer_b200.renderer import make_render_config, make_tile_render_config, tile_app_config
    app = AppConfig.default()
    W, H = 3840, 2160
    full = make_render_config(W, H, app)
    for world in (1, 2, 4, 8):
        ta = tile_app_config(app, world)
        assert ta.train_batch_size * world == app.train_batch_size and ta.train_batch_count == app.train_batch_count
        cfgs = [make_tile_render_config(W, H, ta, r, world) for r in range(world)]
        assert cfgs[0].x_begin == 0 and cfgs[-1].x_end == W
        for a, b in zip(cfgs, cfgs[1:]):
            assert a.x_end == b.x_begin                                              # strips tile the frame without gaps
        assert sum(c.train_width for c in cfgs) == full.train_width                     # lattice columns are partitioned
        for r, c in enumerate(cfgs):
            assert c.train_tx0 == r * c.train_width and c.train_height == full.train_height and c.train_x_dist == full.train_x_dist
            first, last = c.train_tx0 * c.train_x_dist, (c.train_tx0 + c.train_width - 1) * c.train_x_dist
            assert c.x_begin <= first and last < c.x_end                             # the rank's lattice pixels lie inside its strip
            assert c.train_width * c.train_height == ta.train_batch_size * ta.train_batch_count
    import pytest
    with pytest.raises(ValueError):
        tile_app_config(app, 3)


def test_bench_clock_sampler_counts_only_rows_inside_the_load_window():
    """bench.py's nvidia-smi sampler: rows are stamped on arrival and only those inside [begin(), end()] are reported (the query process
    is started before the warm-up); throttle reasons are collected from the window only; no process -> an explicit 'unavailable'."""
    import importlib.util
    import sys
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    sys.modules["bench_under_test"] = bench
    spec.loader.exec_module(bench)
    s = bench.ClockSampler(0)
    assert s.stop()["reasons"] == ["nvidia-smi unavailable"]

    class FakeProc:
        def terminate(self):
            pass

        def poll(self):
            return None
    s = bench.ClockSampler(0)
    s.proc = FakeProc()
    row = lambda sm, cap: [str(sm), "1965", "700.0", "Not Active", "Not Active", "Not Active", cap]
    s.rows = [(1.0, row(300, "Active")), (2.0, row(1965, "Not Active")), (2.5, row(1950, "Not Active")), (3.0, row(1965, "Not Active")), (9.0, row(210, "Active"))]
    s.t0, s.t1 = 1.5, 3.5
    out = s.stop()
    assert out["samples"] == 3 and out["sm_mhz"] == 1965.0 and out["sm_max_mhz"] == 1965.0 and out["reasons"] == []
    s.rows.append((3.2, row(1500, "Active")))
    s.t1 = 3.5
    out = s.stop()
    assert out["samples"] == 4 and out["reasons"] == ["sw_power_cap"]
