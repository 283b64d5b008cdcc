"""the reference's frame metrics (Reference::Result) and its on-disk frame format (RGBA32F OpenEXR).
CPU part: EXR reader / writer against an image written by an independent implementation (OpenCV, fixture
tests/golden/tiny_opencv_rgba32f.exr + .npy, generated with cv2.imwrite(..., IMWRITE_EXR_TYPE_FLOAT)), against the reference's own
bundled frame when the checkout is present, and Result's derived quantities.  GPU part: hpm_compare_images vs the numpy oracle."""
import os

import numpy as np
import pytest

from conftest import ROOT
from nrc_hpm_renderer_b200 import exr
from nrc_hpm_renderer_b200.reference import Result

GOLD = os.path.join(ROOT, "tests", "golden")


def test_read_exr_written_by_opencv():
    img = exr.read_exr(os.path.join(GOLD, "tiny_opencv_rgba32f.exr"))
    assert np.array_equal(img, np.load(os.path.join(GOLD, "tiny_opencv_rgba32f.npy")))       # bit-exact


@pytest.mark.parametrize("comp", [exr.NO_COMPRESSION, exr.ZIPS_COMPRESSION, exr.ZIP_COMPRESSION])
@pytest.mark.parametrize("shape", [(1, 1), (16, 32), (37, 19), (33, 64)])                     # ragged last ZIP block, odd sizes
def test_exr_round_trip(tmp_path, comp, shape):
    rng = np.random.default_rng(shape[0] * 100 + shape[1])
    img = rng.standard_normal((shape[0], shape[1], 4)).astype(np.float32)
    img[0, 0] = [np.inf, -0.0, 1e-40, 0.0]                                                   # inf, signed zero, denormal survive
    p = str(tmp_path / "t.exr")
    exr.write_exr(p, img, comp)
    back = exr.read_exr(p)
    assert back.dtype == np.float32 and np.array_equal(back.view(np.uint32), img.view(np.uint32))


def test_written_header_is_the_tinyexr_flavour(tmp_path):
    """same attributes as the bundled reference/<scene>/0.exr: channels A,B,G,R FLOAT, ZIP, increasing Y"""
    p = str(tmp_path / "t.exr")
    exr.write_exr(p, np.zeros((20, 8, 4), np.float32))
    attrs, _ = exr._read_header(open(p, "rb").read())
    assert [c for c in exr._channels(attrs["channels"][1])] == [("A", 2), ("B", 2), ("G", 2), ("R", 2)]
    assert attrs["compression"][1] == b"\x03" and attrs["lineOrder"][1] == b"\x00"
    assert np.frombuffer(attrs["dataWindow"][1], np.int32).tolist() == [0, 0, 7, 19]


def test_rejects_non_exr(tmp_path):
    p = tmp_path / "x.exr"
    p.write_bytes(b"not an exr file at all")
    with pytest.raises(exr.ExrFormatError):
        exr.read_exr(str(p))


@pytest.mark.skipif(not os.path.exists("/root/reference/reference/0/0.exr"), reason="reference checkout not present")
def test_reads_the_reference_frame_like_the_block_fixture():
    """tests/golden/exr_block8.npz holds 8x8 block means of the reference's frames (read with OpenCV when it was made)"""
    img = exr.read_exr("/root/reference/reference/0/0.exr")
    assert img.shape == (1080, 1920, 4)
    ref = np.load(os.path.join(GOLD, "exr_block8.npz"))["s0"].astype(np.float32)             # [135][240][radiance, alpha], fp16
    blocks = img.reshape(135, 8, 240, 8, 4).mean((1, 3))
    assert np.allclose(blocks[..., 0], ref[..., 0], rtol=2e-3, atol=1e-4) and np.allclose(blocks[..., 3], ref[..., 1], rtol=2e-3, atol=1e-4)


def test_result_derived_quantities():
    r = Result(mse=0.02, refMean=0.5, ownMean=0.45, ownVar=0.09, validPixelCount=10)        # Reference.cpp:10-28
    assert r.GetBias() == pytest.approx(-0.05) and r.GetRelBias() == pytest.approx(-0.1)
    assert r.GetRelVar() == pytest.approx(0.18) and r.GetCV() == pytest.approx(0.3 / 0.45)


def test_oracle_compare_images_known_answer():
    import oracle as O
    ref = np.zeros((2, 2, 4), np.float32); cmp_ = np.zeros((2, 2, 4), np.float32)
    ref[0, 0] = [1, 2, 3, 1]; cmp_[0, 0] = [2, 2, 2, 1]
    ref[0, 1] = [0, 0, 0, 0.5]; cmp_[0, 1] = [3, 3, 3, 0]
    ref[1, 0] = [9, 9, 9, 0]; cmp_[1, 0] = [100, 100, 100, 1]                                # ref alpha 0: ignored
    r = O.compare_images(ref, cmp_)
    assert r["validPixelCount"] == 2
    assert r["mse"] == pytest.approx(((1 + 0 + 1) / 3 + 27 / 3) / 2)
    assert r["refMean"] == pytest.approx(1.0) and r["ownMean"] == pytest.approx(2.5)
    assert r["ownVar"] == pytest.approx((3 * 0.25 / 3 + 3 * 0.25 / 3) / 2)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1080, 1920), (64, 96), (3, 5)])
def test_compare_images_kernel_vs_oracle(shape):
    import torch
    import oracle as O
    from nrc_hpm_renderer_b200.reference import compare_device_images
    rng = np.random.default_rng(5)
    h, w = shape
    ref = rng.random((h, w, 4), dtype=np.float32); cmp_ = (ref + 0.1 * rng.standard_normal((h, w, 4))).astype(np.float32)
    ref[..., 3] = (rng.random((h, w)) < 0.3).astype(np.float32)
    d_ref, d_cmp = torch.from_numpy(ref).cuda(), torch.from_numpy(cmp_).cuda()
    a = compare_device_images(d_ref.data_ptr(), d_cmp.data_ptr(), w, h)
    b = compare_device_images(d_ref.data_ptr(), d_cmp.data_ptr(), w, h)
    assert a == b                                                                             # deterministic
    o = O.compare_images(ref, cmp_)
    assert a.validPixelCount == o["validPixelCount"]                                          # integer: exact
    for k in ("mse", "refMean", "ownMean", "ownVar"):
        assert getattr(a, k) == pytest.approx(o[k], rel=2e-6), k                              # fp32 per-pixel terms, fp64 sums
    none = torch.zeros_like(d_ref)
    z = compare_device_images(none.data_ptr(), d_cmp.data_ptr(), w, h)                        # no valid pixel: all zero, no NaN
    assert z.validPixelCount == 0 and z.mse == 0 and z.ownVar == 0


@pytest.mark.gpu
def test_reference_class_round_trip(tmp_path):
    """Reference generates reference/<id>/0.exr with the MC renderer when the folder is missing, reloads it, and CompareMc of the
    same converged renderer state against it gives a small error; ExportOutputImageToFile writes what GetImage returns."""
    from nrc_hpm_renderer_b200 import AppConfig, Camera, HpmSceneConfig
    from nrc_hpm_renderer_b200 import renderer as R
    from nrc_hpm_renderer_b200.reference import Reference
    grid = np.ascontiguousarray(np.load(os.path.join(GOLD, "wdas_cloud_sixteenth_u8.npz"))["data"])
    app = AppConfig.default(); app.scene = HpmSceneConfig.preset(0)
    scene = R.HpmScene(grid, app.scene)
    W, H = 96, 64
    ref = Reference(W, H, app, scene, reference_root=str(tmp_path / "reference"), frames=256, path_length=8)
    assert os.path.exists(tmp_path / "reference" / "0" / "0.exr")
    again = Reference(W, H, app, scene, reference_root=str(tmp_path / "reference"))          # loads, does not regenerate
    a, b = again.m_RefImage.cpu().numpy(), ref.m_RefImage.cpu().numpy()
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert np.isfinite(a).all(), np.argwhere(~np.isfinite(a))[:8]
    cam = Camera(aspect=W / H, pos=(60.0, 5.0, 0.0))
    mc = R.McHpmRenderer(W, H, 8, True, cam, scene)
    rng = np.random.default_rng(11)
    res = None
    for _ in range(128):
        res = ref.CompareMc(mc, cam, rng.random(4).astype(np.float32))
    assert mc.camera is cam                                                                   # camera restored
    assert res.validPixelCount == int((ref.m_RefImage[..., 3] != 0).sum())
    assert np.isfinite([res.mse, res.ownVar]).all() and abs(res.GetRelBias()) < 0.1        # same estimator, same camera: unbiased
    p = str(tmp_path / "out.exr")
    mc.ExportOutputImageToFile(p)
    assert np.array_equal(exr.read_exr(p), mc.GetImage())
