"""GPU parity of the NRC kernels (through the C ABI) against
  (1) the committed tcnn fixtures (tests/golden/tcnn_*.npz: outputs of the reference's own tiny-cuda-nn on a B200), and
  (2) the CPU oracle (oracle/nrc_oracle.cpp) on the same seeded inputs.
Tolerance (BASELINE north_star: max relative error <= 1e-2 per output): err = |a-b| / max(|b|, rms(b)) <= 1e-2, i.e. every
element within 1 % of its own magnitude or, for elements below the tensor's rms, within 1 % of the rms.  Activations are
stored in fp16 between layers (ulp 5e-4 relative) and tcnn additionally accumulates in fp16, so outputs much
smaller than the rms (cancellation) cannot agree to 1 % of their own magnitude between ANY two implementations; gradients
(tcnn: fp16 split-K accumulation over the batch) get the stated wider bound."""
import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu

CONFIGS = ["hash_ob_d6", "tri_ob_d5", "hash_tri_d3", "id_id_d2", "freq_ob_d4",
           "hash_ob_d6_w128", "tri_ob_d5_w128"]          # nnWidth = 128: the 128-neuron kernels (csrc/nrc_wide_kernels.cuh)


def rel_err(a, b, floor=1.0):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    ok = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), ok)
    scale = np.maximum(np.abs(b[ok]), floor * np.sqrt(np.mean(b[ok] ** 2)) + 1e-30)
    return float(np.max(np.abs(a[ok] - b[ok]) / scale)) if ok.any() else 0.0


def contract_distribution(a, b, floor=1e-2):
    """error distribution under the contract metric of SURVEY.md 8(d): |a-b| / max(|b|, 1e-2 rms(b)) per output element"""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    ok = np.isfinite(b)
    e = np.abs(a[ok] - b[ok]) / np.maximum(np.abs(b[ok]), floor * np.sqrt(np.mean(b[ok] ** 2)) + 1e-30)
    return {"p50": float(np.percentile(e, 50)), "p99": float(np.percentile(e, 99)), "p99.9": float(np.percentile(e, 99.9)), "max": float(e.max()),
            "frac_gt_1e-2": float((e > 1e-2).mean())}


def report_parity(name, what, ours, baseline):
    """append the measured distribution to gpurun_out/parity_contract.jsonl (copied to profiles/ by the builder)"""
    import json, os
    from conftest import ROOT
    d = os.path.join(ROOT, "gpurun_out")
    line = {"config": name, "what": what, "metric": "|a-b| / max(|b|, 1e-2 rms(b))", "cuda_path_vs_tcnn": ours, "oracle_fp32_accumulation_vs_tcnn": baseline}
    print(json.dumps(line))
    if os.path.isdir(d):
        with open(os.path.join(d, "parity_contract.jsonl"), "a") as f:
            f.write(json.dumps(line) + "\n")


def make_cache(z, **kw):
    import torch  # noqa: F401
    from nrc_hpm_renderer_b200 import AppConfig
    from nrc_hpm_renderer_b200.nrc import NeuralRadianceCache
    app = AppConfig.default()
    app.pos_enc_id, app.dir_enc_id, app.nn_depth = int(z["pos"]), int(z["dir"]), int(z["depth"])
    app.nn_width = int(z["width"]) if "width" in z else 64
    return NeuralRadianceCache(app, **kw)


def dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("name", CONFIGS)
def test_init_params_bit_exact(name):
    from nrc_hpm_renderer_b200 import nrc as N
    z = golden(f"tcnn_{name}.npz")
    c = make_cache(z)
    assert c.n_params == int(z["n_params"]) and c.n_mlp_params == int(z["n_mlp"]) and c.input_width == int(z["padded_input"])
    master = c.get_params(N.MASTER)
    assert np.array_equal(master[: c.n_mlp_params], z["params_init_mlp"])
    if c.n_params > c.n_mlp_params:
        assert np.array_equal(master[z["grid_idx"]], z["params_init_grid"])
    assert abs(master.astype(np.float64).sum() - float(z["params_init_sum"])) < 1e-9
    assert np.all(c.get_params(N.EMA) == 0)          # Q7: EMA weights start at zero


@pytest.mark.parametrize("name", CONFIGS)
def test_encoding_vs_tcnn(name):
    import torch
    z = golden(f"tcnn_{name}.npz")
    c = make_cache(z)
    n = int(z["n_infer"])
    out = torch.zeros((n, c.input_width), dtype=torch.float16, device="cuda")
    c.encode(dev(z["infer_in"]), n, out, use_ema=False)
    torch.cuda.synchronize()
    got = out.cpu().numpy().astype(np.float32); ref = z["network_input"].astype(np.float32)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    d = np.abs(got - ref); d[np.isnan(d)] = 0
    assert d.max() <= 2e-3, d.max()                  # fp16 ulp at |x|<=2 is 1e-3..2e-3 (fast-math differences only)
    assert np.mean(d > 0) < 0.02


@pytest.mark.parametrize("name", CONFIGS)
def test_inference_vs_tcnn(name):
    import torch
    z = golden(f"tcnn_{name}.npz")
    c = make_cache(z)
    n = int(z["n_infer"])
    out = torch.full((n, 3), 7.0, dtype=torch.float32, device="cuda")
    c.inference(dev(z["infer_in"]), out, n, use_ema=True)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), z["infer_ema_step0"])     # frame 0: EMA weights are zero -> exact zeros (Q7)
    c.inference(dev(z["infer_in"]), out, n, use_ema=False)
    torch.cuda.synchronize()
    assert rel_err(out.cpu().numpy(), z["infer_working_step0"], floor=1.0) <= 1e-2
    # ---- the same outputs under SURVEY.md 8(d)'s contract metric |a-b| / max(|b|, 1e-2 rms): reported as a distribution, and
    # bounded by what a FAITHFUL restatement with the CUDA path's arithmetic (fp32 accumulation) reaches against the same fixture.
    # Measured on the CPU (oracle vs these fixtures): even in tcnn's own fp16-accumulation mode, where the oracle reproduces
    # 99.9 % of tcnn's outputs bit for bit, the maximum is 1.2e-2 (hash_ob_d6) / 3.1e-2 (tri_ob_d5); in fp32 mode p99 is 8e-3 / 7.8e-2:
    # outputs two decades below the rms are differences of O(rms) terms rounded to fp16 (ulp 5e-4 relative) layer by layer.
    import oracle as O
    o = O.NrcOracle(O.nrc_config(int(z["pos"]), int(z["dir"]), int(z["depth"]), int(z["width"]) if "width" in z else 64, accum_fp16=0))
    ref = z["infer_working_step0"]
    dist = contract_distribution(out.cpu().numpy(), ref)
    base = contract_distribution(o.inference(z["infer_in"], use_ema=False), ref)
    report_parity(name, "inference_vs_tcnn", dist, base)
    if name != "freq_ob_d4":          # Frequency: tcnn evaluates __sinf (frequency.h:74) on arguments up to 2^11 pi; covered by the rms-floored bound above
        assert dist["p50"] <= (3e-3 if "w128" in name else 2.5e-3)      # 128 neurons: twice as many fp16 products per accumulator
        assert dist["p99"] <= max(1e-2, 1.5 * base["p99"]), (dist, base)
        assert dist["frac_gt_1e-2"] <= base["frac_gt_1e-2"] + 0.02, (dist, base)


@pytest.mark.parametrize("name", CONFIGS)
def test_training_vs_tcnn(name):
    import torch
    from nrc_hpm_renderer_b200 import nrc as N
    z = golden(f"tcnn_{name}.npz")
    c = make_cache(z)
    B, steps, nm = int(z["batch"]), int(z["steps"]), int(z["n_mlp"])
    tin, tgt = dev(z["train_in"]), dev(z["train_tgt"])
    losses = []
    for s in range(steps):
        c.training_step(tin[s * B:(s + 1) * B], tgt[s * B:(s + 1) * B], B, run_optimizer=(s > 0))
        if s == 0:
            torch.cuda.synchronize()
            out16 = c.last_step_tensor(0, B); dout = c.last_step_tensor(1, B)
            assert rel_err(out16[:, :3], z["output_step0"].astype(np.float32)[:, :3], floor=1.0) <= 1e-2
            assert rel_err(dout[:, :3], z["dL_doutput_step0"].astype(np.float32)[:, :3], floor=1.0) <= 2e-2
            assert np.all(dout[:, 3:] == 0)
            g = c.get_params(N.GRAD)       # before the optimizer consumes (and re-zeroes) the encoding gradient
            c.optimizer_step()
            # MLP weight gradients: fp32 accumulation here vs fp16 accumulation in tcnn (Q8)
            assert rel_err(g[:nm], z["grad0_mlp"].astype(np.float32), floor=1.0) <= 1e-1
            # Adam's first step is lr * g / (|g| + eps): its size is ~lr whatever |g| is, so a gradient entry whose sign is
            # not stable under fp16-vs-fp32 accumulation moves the weight by up to 2*lr the other way.  Entries with a
            # clearly non-zero reference gradient must agree to tolerance; every entry must stay within 2*lr.
            lr = 0.01
            gref = z["grad0_mlp"].astype(np.float32)
            fin = np.isfinite(gref)       # NaN inputs (Q5 with a non-OneBlob direction encoding) poison dW0 in tcnn and here alike
            assert np.array_equal(np.isfinite(g[:nm]), fin)
            sure = fin & (np.abs(np.nan_to_num(gref)) > 0.05 * np.sqrt(np.mean(gref[fin] ** 2)))
            p1 = c.get_params(N.MASTER); e1 = c.get_params(N.EMA)
            assert rel_err(p1[:nm][sure], z["params_step1_mlp"][sure]) <= 1e-2
            assert np.nanmax(np.abs(p1[:nm] - z["params_step1_mlp"])) <= 2.05 * lr
            assert rel_err(e1[:nm][sure], z["ema_step1_mlp"].astype(np.float32)[sure]) <= 1e-2
            if c.n_params > nm:
                gi = z["grid_idx"]
                gg = z["grad0_grid"].astype(np.float32)
                assert np.array_equal(g[gi] != 0, gg != 0) or np.mean((g[gi] != 0) != (gg != 0)) < 0.02   # same touched entries (fp16 underflow aside)
                # tcnn computes dL/dinput with fp16 accumulators (cutlass_matmul.h:67-68) and scatters with
                # order-dependent fp16x2 atomics (grid.h:252-255).  tests/test_oracle_nrc.py pins the size of that effect on
                # the CPU: the oracle in fp16-accumulation mode matches this fixture to 6e-4 relative L2, in fp32 mode
                # (the arithmetic of the CUDA path) it differs by 4.2e-2.  The CUDA path must sit at the fp32 figure.
                d = np.abs(g[gi].astype(np.float64) - gg)
                assert np.linalg.norm(d) / np.linalg.norm(gg) <= 6e-2
                sure_g = np.abs(gg) > 0.05 * np.sqrt(np.mean(gg[gg != 0] ** 2))
                stable = sure_g & (d <= 0.1 * np.abs(gg))
                assert rel_err(p1[gi][stable], z["params_step1_grid"][stable]) <= 2e-2
                assert np.max(np.abs(p1[gi] - z["params_step1_grid"])) <= 2.05 * lr
        losses.append(c.GetLoss())
    losses = np.array(losses); ref = z["losses"]
    assert np.all(np.isfinite(losses))
    assert abs(losses[0] - ref[0]) / ref[0] <= 1e-3
    # later steps see slightly different weights (fp16 vs fp32 accumulation, atomic order): curve must track.  The 128-neuron
    # fixtures contain a loss SPIKE (lr 0.01 on random targets: x2.3 at step 7 / x13 at step 12 in tcnn's own curve); from there on
    # the trajectory is chaotic in the arithmetic -- the CPU oracle reproduces tcnn's curve to 0.3 % in fp16-accumulation mode and
    # leaves it by 12 % at the spike in fp32 mode (tests/test_oracle_nrc.py) -- so the bound against tcnn holds up to the spike and
    # the rest of the curve is held against the oracle with the CUDA path's own arithmetic (fp32 accumulation).
    spike = next((i for i in range(2, len(ref)) if ref[i] > 1.5 * ref[i - 1]), len(ref))
    assert np.max(np.abs(losses - ref)[:spike] / np.maximum(ref, 1e-3)[:spike]) <= 0.08, (losses, ref)
    out = torch.zeros((int(z["n_infer"]), 3), dtype=torch.float32, device="cuda")
    c.inference(dev(z["infer_in"]), out, int(z["n_infer"]), use_ema=True)
    torch.cuda.synchronize()
    if spike == len(ref):
        ref_out = z["infer_ema_final"]
    else:
        import oracle as O
        o = O.NrcOracle(O.nrc_config(int(z["pos"]), int(z["dir"]), int(z["depth"]), int(z["width"]) if "width" in z else 64, accum_fp16=0))
        lo = np.array([o.training_step(z["train_in"][s * B:(s + 1) * B], z["train_tgt"][s * B:(s + 1) * B]) for s in range(steps)])
        assert np.max(np.abs(losses - lo) / np.maximum(lo, 1e-3)) <= 0.15, (losses, lo)
        ref_out = o.inference(z["infer_in"], use_ema=True)
    err = np.abs(out.cpu().numpy() - ref_out)
    assert np.nanmax(err) <= (0.05 if spike == len(ref) else 0.15) * max(1.0, np.sqrt(np.nanmean(ref_out ** 2)))      # same bound as the post-spike losses


@pytest.mark.parametrize("pos,dr,depth,width", [(0, 0, 5, 64), (2, 0, 6, 64), (3, 0, 4, 64), (1, 2, 1, 64), (0, 1, 8, 64), (3, 2, 8, 64),
                                                 (0, 0, 6, 128), (2, 0, 5, 128), (1, 1, 1, 128), (0, 2, 3, 128)])
def test_against_oracle(pos, dr, depth, width, oracle_lib):
    """same seeded inputs through the CUDA path and the CPU oracle: inference, one training step, gradients"""
    import torch
    from nrc_hpm_renderer_b200 import AppConfig, nrc as N
    O = oracle_lib
    app = AppConfig.default(); app.pos_enc_id, app.dir_enc_id, app.nn_depth, app.nn_width = pos, dr, depth, width
    c = N.NeuralRadianceCache(app)
    o = O.NrcOracle(O.nrc_config(pos, dr, depth, width))
    assert c.n_params == o.n_params
    assert np.array_equal(c.get_params(N.MASTER), o.get(o.MASTER))
    rng = np.random.default_rng(5 + pos * 7 + dr)
    n = 1000                                            # ragged: not a multiple of the 128-record tile
    rec = rng.random((n, 5), dtype=np.float32)
    if pos != 3:     # Frequency: tcnn uses __sinf (frequency.h:74), whose error grows with |x|; the oracle's sinf only tracks it for x in [0,1)
        rec[: n // 2, :3] += np.array([31.1585, 21.1475, 38.3535], np.float32)
    rec[:, 3] = rec[:, 3] * 2 - 0.5
    rec[rng.random(n) < 0.1, 4] = np.nan
    out = torch.zeros((n, 3), dtype=torch.float32, device="cuda")
    c.inference(dev(rec), out, n, use_ema=False)
    torch.cuda.synchronize()
    tol = 1e-1 if pos == 3 else 1e-2
    assert rel_err(out.cpu().numpy(), o.inference(rec, use_ema=False)) <= tol
    B = 256
    tin = rec[:B].copy(); tgt = (rng.random((B, 3), dtype=np.float32) * 2).astype(np.float32)
    c.training_step(dev(tin), dev(tgt), B, run_optimizer=False)
    lo = o.training_step(tin, tgt, run_optimizer=False)
    assert abs(c.GetLoss() - lo) / lo <= (5e-2 if pos == 3 else 2e-3)
    if pos != 3:
        g = c.get_params(N.GRAD); go = o.get(o.GRAD)
        nm = c.n_mlp_params
        assert rel_err(g[:nm], go[:nm]) <= 2e-2
        if c.n_params > nm:
            # hash-grid gradient: same touched entries; values differ only by the fp16 atomic accumulation order
            touched = go[nm:] != 0
            assert np.mean((g[nm:] != 0) != touched) < 1e-3
            big = np.abs(go[nm:]) > 0.05 * np.sqrt(np.mean(go[nm:][touched] ** 2))
            assert rel_err(g[nm:][big], go[nm:][big]) <= 5e-2
        c.optimizer_step()
        assert np.all(c.get_params(N.GRAD)[nm:] == 0)


def test_indexed_inference_matches_dense():
    import torch
    from nrc_hpm_renderer_b200 import AppConfig, nrc as N
    c = N.NeuralRadianceCache(AppConfig.default())
    c.set_ema(c.get_params(N.MASTER))
    rng = np.random.default_rng(3)
    n = 4096
    rec = dev(rng.random((n, 5), dtype=np.float32))
    dense = torch.zeros((n, 3), dtype=torch.float32, device="cuda")
    c.inference(rec, dense, n)
    idx = np.sort(rng.choice(n, 1500, replace=False)).astype(np.uint32)
    sparse = torch.full((n, 3), -1.0, dtype=torch.float32, device="cuda")
    cnt = torch.tensor([len(idx)], dtype=torch.int32, device="cuda")
    c.inference_indexed(rec, sparse, dev(idx.view(np.int32)), cnt, n)
    torch.cuda.synchronize()
    d, s = dense.cpu().numpy(), sparse.cpu().numpy()
    assert np.array_equal(s[idx], d[idx])
    mask = np.ones(n, bool); mask[idx] = False
    assert np.all(s[mask] == -1.0)


@pytest.mark.parametrize("width", [64, 128])
def test_empty_and_ragged_batches(width):
    """Edge sizes of Inference(): no records, one record, one short of / one past a 128-record tile, a ragged multi-tile batch, and an
    empty compaction list.  A record's result must not depend on the batch it sits in (bit for bit against the same rows of one large
    batch), rows past the batch must stay untouched, and n = 0 must be a no-op instead of an error
    (reference: src/NeuralRadianceCache.cu:100-118 loops over however many batches the filter leaves, possibly none)."""
    import torch
    from nrc_hpm_renderer_b200 import AppConfig, nrc as N
    app = AppConfig.default(); app.nn_width = width
    c = N.NeuralRadianceCache(app)
    c.set_ema(c.get_params(N.MASTER))
    rng = np.random.default_rng(11)
    n_all = 148 * 128 * 2 + 77
    rec = dev(rng.random((n_all, 5), dtype=np.float32))
    full = torch.zeros((n_all, 3), dtype=torch.float32, device="cuda")
    c.inference(rec, full, n_all)
    torch.cuda.synchronize()
    ref = full.cpu().numpy()
    assert np.any(ref != 0)
    for n in (0, 1, 127, 128, 129, 1000, 148 * 128 + 1):
        out = torch.full((n_all, 3), -7.0, dtype=torch.float32, device="cuda")
        c.inference(rec, out, n)
        torch.cuda.synchronize()
        o = out.cpu().numpy()
        assert np.array_equal(o[:n], ref[:n]), n
        assert np.all(o[n:] == -7.0), n
    # compacted inference with an empty list: nothing is written
    out = torch.full((n_all, 3), -7.0, dtype=torch.float32, device="cuda")
    idx = torch.zeros(16, dtype=torch.int32, device="cuda"); cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    c.inference_indexed(rec, out, idx, cnt, n_all)
    torch.cuda.synchronize()
    assert np.all(out.cpu().numpy() == -7.0)
    # host entry point with no records
    assert c.inference_host(np.zeros((0, 5), np.float32), use_ema=True).shape == (0, 3)


def test_host_entry_points_and_errors():
    from nrc_hpm_renderer_b200 import AppConfig, nrc as N, _lib
    c = N.NeuralRadianceCache(AppConfig.default())
    rng = np.random.default_rng(0)
    rec = rng.random((512, 5), dtype=np.float32)
    out = c.inference_host(rec, use_ema=True)
    assert out.shape == (512, 3) and np.all(out == 0)
    loss = c.training_step_host(rec[:256], rng.random((256, 3), dtype=np.float32))
    assert np.isfinite(loss) and loss > 0
    with pytest.raises(_lib.NrcHpmError):
        c.training_step_host(rec[:100], rng.random((100, 3), dtype=np.float32))      # not a multiple of 128
    with pytest.raises(_lib.NrcHpmError):
        N.NeuralRadianceCache(config_json={"encoding": {"otype": "SphericalHarmonics"}})


@pytest.mark.parametrize("n", [700_000, 1_300_001])          # uniform chunks / graded chunks (large first chunk, small last ones), ragged
def test_infer_and_train_host_equals_separate_calls(n):
    """nrc_infer_and_train_host (one pipelined call per frame) == nrc_inference_host followed by nrc_training_step_host per batch:
    identical radiance (inference is deterministic and reads the pre-training EMA weights) and the same loss trajectory (the
    forward pass and the loss are deterministic; the fp16 atomics of the grid gradient make later losses agree to tolerance)."""
    from nrc_hpm_renderer_b200 import AppConfig, nrc as N
    rng = np.random.default_rng(5)
    B, nb = 1024, 3                                   # several pipeline chunks, ragged last chunk
    rec = rng.random((n, 5), dtype=np.float32)
    tin = rng.random((B * nb, 5), dtype=np.float32)
    tgt = (rng.random((B * nb, 3), dtype=np.float32) * 2).astype(np.float32)
    a, b = N.NeuralRadianceCache(AppConfig.default()), N.NeuralRadianceCache(AppConfig.default())
    out_a, out_b = np.full((n, 3), -1, np.float32), None
    for frame in range(3):
        la = a.infer_and_train_host(rec, out_a, tin, tgt, B, True)
        out_b = b.inference_host(rec, use_ema=True)
        for k in range(nb):
            lb = b.training_step_host(tin[k * B:(k + 1) * B], tgt[k * B:(k + 1) * B])
        assert np.array_equal(out_a, out_b) if frame == 0 else rel_err(out_a, out_b) <= 1e-2
        assert abs(la - lb) <= 2e-3 * abs(lb)
        assert a.GetLoss() == la
    assert np.any(out_a != 0)                        # the EMA weights moved after the first frame
    # halves can be skipped
    assert np.isfinite(a.infer_and_train_host(None, None, tin, tgt, B, True))
    a.infer_and_train_host(rec[:4096], out_a[:4096], None, None, 0, True)
