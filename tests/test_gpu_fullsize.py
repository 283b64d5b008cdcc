"""GPU: the path at BASELINE.json's FULL sizes (1920x1080: 2 073 600 inference records, 4 x 2^14 training records), checked through
size-independent properties -- the oracle finishes only small cases in seconds, so at full size the kernels are checked against
themselves (launch-shape / chunking / permutation / compaction invariance, bit for bit), against the oracle on a random sample,
and through conservation laws of the frame (every scattered pixel is evaluated exactly once, frame 0 shows the primary estimate)."""
import os

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

W, H = 1920, 1080
N = W * H
SKY_HALF = np.array([62.317, 42.295, 76.707], np.float32) / 2


def records(rng, n):
    rec = rng.random((n, 5), dtype=np.float32)
    rec[:, :3] += SKY_HALF
    rec[:, 3] = rec[:, 3] * 2 - 0.5
    return rec


@pytest.fixture(scope="module")
def trained_cache():
    import torch
    from nrc_hpm_renderer_b200 import AppConfig, nrc as NR
    c = NR.NeuralRadianceCache(AppConfig.default())
    rng = np.random.default_rng(5)
    for _ in range(4):                                            # non-trivial weights and EMA
        rec = torch.from_numpy(records(rng, 1 << 14)).cuda(); tgt = torch.from_numpy((rng.random((1 << 14, 3), dtype=np.float32) * 2).astype(np.float32)).cuda()
        c.training_step(rec, tgt, 1 << 14, True)
    torch.cuda.synchronize()
    return c


def test_full_frame_inference_invariances(trained_cache, oracle_lib):
    import torch
    from nrc_hpm_renderer_b200 import nrc as NR
    c = trained_cache
    rng = np.random.default_rng(17)
    rec = records(rng, N)
    rec[::97, 4] = np.nan                                          # Q5: the reference produces NaN phi for a share of the records
    d_rec = torch.from_numpy(rec).cuda()
    full = torch.zeros((N, 3), dtype=torch.float32, device="cuda")
    c.inference(d_rec, full, N)                                    # one persistent launch (the reference's single 2^21-capped batch)
    torch.cuda.synchronize()
    ref = full.cpu().numpy()
    assert np.isfinite(ref).all()                                  # tcnn's ReLU maps the NaN features to 0 (Q5)
    assert np.abs(ref).max() > 0
    # (1) chunking / launch-shape invariance: ragged chunks (1 record, odd sizes, a single-warpgroup launch) give the same bits
    parts = torch.zeros_like(full)
    edges = [0, 1, 130, 4096 + 37, 300_001, 1_000_003, N]
    for a, b in zip(edges, edges[1:]):
        c.inference(d_rec[a:b], parts[a:b], b - a)
    c.inference(d_rec, parts, 0)                                   # n = 0 is a no-op
    torch.cuda.synchronize()
    assert torch.equal(parts, full)
    # (2) permutation equivariance
    perm = rng.permutation(N)
    out_p = torch.zeros_like(full)
    c.inference(torch.from_numpy(rec[perm]).cuda(), out_p, N)
    torch.cuda.synchronize()
    assert np.array_equal(out_p.cpu().numpy(), ref[perm])
    # (3) compaction with the identity list == dense; with a random third of the records == dense on those, untouched elsewhere
    idx_all = torch.arange(N, dtype=torch.int32, device="cuda"); cnt = torch.tensor([N], dtype=torch.int32, device="cuda")
    out_i = torch.zeros_like(full)
    c.inference_indexed(d_rec, out_i, idx_all, cnt, N)
    sel = np.sort(rng.choice(N, N // 3, replace=False)).astype(np.int32)
    out_s = torch.full((N, 3), -7.0, dtype=torch.float32, device="cuda")
    c.inference_indexed(d_rec, out_s, torch.from_numpy(rng.permutation(sel)).cuda(), torch.tensor([len(sel)], dtype=torch.int32, device="cuda"), N)
    torch.cuda.synchronize()
    assert torch.equal(out_i, full)
    s = out_s.cpu().numpy()
    mask = np.zeros(N, bool); mask[sel] = True
    assert np.array_equal(s[mask], ref[mask]) and np.all(s[~mask] == -7.0)
    # (4) the snapshot evaluates the captured parameters
    c.snapshot_params(True)
    out_n = torch.zeros_like(full)
    c.inference(d_rec, out_n, N, NR.SNAPSHOT)
    torch.cuda.synchronize()
    assert torch.equal(out_n, full)
    # (5) a random sample against the CPU oracle loaded with the same parameters (tolerance: 1e-2 of max(|ref|, rms), fp16 network)
    o = oracle_lib.NrcOracle(oracle_lib.nrc_config(0, 0, 6))
    o.set_params(c.get_params(NR.MASTER)); o.set_ema(c.get_params(NR.EMA))
    pick = rng.choice(N, 4096, replace=False)
    want = o.inference(rec[pick], use_ema=True)
    scale = np.maximum(np.abs(want), np.sqrt(np.mean(want ** 2)))
    assert np.max(np.abs(ref[pick] - want) / scale) <= 1e-2


def test_full_size_training_is_reproducible():
    """MLP-only configuration (no atomics anywhere): the four 2^14-record steps of a frame give bit-identical losses and parameters
    on two caches, and through the one-call host path."""
    import torch
    from nrc_hpm_renderer_b200 import AppConfig, nrc as NR
    rng = np.random.default_rng(23)
    B, K = 1 << 14, 4
    rec = records(rng, B * K); tgt = (rng.random((B * K, 3), dtype=np.float32) * 2).astype(np.float32)
    runs = []
    for mode in ("device", "device", "host"):
        app = AppConfig.default(); app.pos_enc_id, app.nn_depth = 2, 6
        c = NR.NeuralRadianceCache(app)
        if mode == "device":
            d_rec, d_tgt = torch.from_numpy(rec).cuda(), torch.from_numpy(tgt).cuda()
            losses = []
            for b in range(K):
                c.training_step(d_rec[b * B:(b + 1) * B], d_tgt[b * B:(b + 1) * B], B, True)
                losses.append(c.GetLoss())
        else:
            out = np.zeros((256, 3), np.float32)
            c.infer_and_train_host(rec[:256], out, rec, tgt, B, True)
            losses = [c.GetLoss()]
        runs.append((losses, c.get_params(NR.MASTER), c.get_params(NR.EMA)))
    assert runs[0][0] == runs[1][0] and np.array_equal(runs[0][1], runs[1][1]) and np.array_equal(runs[0][2], runs[1][2])
    assert runs[2][0][-1] == runs[0][0][-1] and np.array_equal(runs[2][1], runs[0][1])
    assert runs[0][0][-1] < runs[0][0][0]                          # and it learns


def test_full_frame_conservation_laws():
    """1080p frame on the bundled cloud: every scattered pixel is in the compaction list exactly once, its record is non-zero and
    every other record is zero, the batch filter flags are the OR of the pixels they cover, and frame 0 (EMA still zero, Q7) shows
    exactly the primary estimate."""
    from nrc_hpm_renderer_b200 import AppConfig, Camera, HpmSceneConfig, volume
    from nrc_hpm_renderer_b200 import renderer as R
    from nrc_hpm_renderer_b200.nrc import NeuralRadianceCache
    path = os.path.join(ROOT, "data", "wdas_cloud_quarter_u8.npz")
    if not os.path.exists(path):
        pytest.skip("bundled volume fixture missing")
    grid = volume.load_volume(path).data
    app = AppConfig.default(); app.scene = HpmSceneConfig.preset(0); app.log2_infer_batch_size = 16
    nrc = NeuralRadianceCache(app)
    scene = R.HpmScene(grid, app.scene)
    r = R.NrcHpmRenderer(W, H, False, Camera(aspect=W / H), app, scene, nrc)
    r.Render(True, np.array([0.11, 0.52, 0.93, 0.34], np.float32)); r.sync()
    info = r.read(R.BUF_PRIMARY_INFO).reshape(H, W)                # pixel (x, y) at y*W + x
    rec = r.read(R.BUF_INFER_INPUT).reshape(-1, 5)                 # record of pixel (x, y) at x*H + y
    scattered = info.T.reshape(-1) == 1.0
    n_sc = int(scattered.sum())
    assert 0.1 * N < n_sc < 0.6 * N
    assert int(r.read(R.BUF_COUNTERS)[2]) == n_sc                  # compaction count == number of scattered pixels
    assert np.all(rec[~scattered] == 0)                            # vkCmdFillBuffer semantics folded into the kernel
    assert np.all(np.abs(rec[scattered, :3]).sum(1) > 0)
    flt = r.read(R.BUF_INFER_FILTER)
    bs = app.infer_batch_size
    want = np.array([scattered[i * bs:(i + 1) * bs].any() for i in range(len(flt))])
    assert np.array_equal(flt != 0, want)
    img = r.GetImage()
    prim = r.read(R.BUF_PRIMARY_COLOR).reshape(H, W, 4)
    assert np.array_equal(img[..., :3], prim[..., :3], equal_nan=True)        # frame 0: no cache term
    # a second frame uses the trained cache: non-negative term on scattered pixels only
    r.Render(True, np.array([0.21, 0.62, 0.03, 0.44], np.float32)); r.sync()
    img2 = r.GetImage(); prim2 = r.read(R.BUF_PRIMARY_COLOR).reshape(H, W, 4); info2 = r.read(R.BUF_PRIMARY_INFO).reshape(H, W)
    d = img2[..., :3] - prim2[..., :3]
    assert np.all(d[info2 != 1.0] == 0) and np.all(d[info2 == 1.0] >= 0) and d.max() > 0
