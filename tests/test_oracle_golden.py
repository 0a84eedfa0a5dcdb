"""Pin oracle/spn_oracle.py (numpy restatement) against the fixtures produced by
running the reference's own modules (tests/golden/make_golden.py)."""
import glob
import os

import numpy as np
import pytest

from oracle import spn_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PP = sorted(glob.glob(os.path.join(GOLDEN, "pp_*.npz")) + glob.glob(os.path.join(GOLDEN, "lrru_*.npz")))
NL = sorted(glob.glob(os.path.join(GOLDEN, "nlspn_*.npz")))
GEN = sorted(glob.glob(os.path.join(GOLDEN, "gen_*.npz")))


def _close(a, b, rtol, atol, what):
    a = np.asarray(a)
    b = np.asarray(b)
    err = np.abs(a - b)
    tol = atol + rtol * np.abs(b)
    assert (err <= tol).all(), f"{what}: max abs err {err.max():.3e}, max |ref| {np.abs(b).max():.3e}"


def test_fixtures_present():
    assert len(PP) == 12 and len(NL) == 5 and len(GEN) == 4


@pytest.mark.parametrize("path", GEN, ids=[os.path.basename(p)[:-4] for p in GEN])
@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_generator_tail_matches_reference(path, prec):
    """Generator.conv_weight / conv_offset / zero-centre insert + PostProcessor, forward and all gradients,
    against the reference's own Generator run (tests/golden/make_golden_generator.py)."""
    z = np.load(path)
    dt = np.float64 if prec == "f64" else np.float32
    rtol, atol = (1e-11, 1e-11) if prec == "f64" else (1e-5, 4e-6)
    feature = z["in_feature"].astype(dt)
    cww, cwb, cow, cob = (z["in_" + k].astype(dt) for k in ("conv_weight_w", "conv_weight_b", "conv_offset_w", "conv_offset_b"))
    weight, offset = O.generator_tail(feature, cww, cwb, cow, cob)
    _close(weight, z[prec + "_weight"], rtol, atol, "weight")
    _close(offset, z[prec + "_offset"], rtol, atol, "offset")
    assert (offset[:, 8:10] == 0).all()
    init, gout = z["in_init"].astype(dt), z["in_grad_out"].astype(dt)
    w9, b1 = z["in_w"].astype(dt).reshape(9), z["in_b"].astype(dt)[0]
    mode, scale = int(z["norm_mode"]), float(z["scale"])
    # the propagation consumes the reference's own weight/offset here so that the two stages are pinned separately
    wr, orf = z[prec + "_weight"].astype(dt), z[prec + "_offset"].astype(dt)
    _close(O.postprocessor_forward(init, wr, orf, w9, b1, mode, scale), z[prec + "_out"], rtol, atol, "out")
    g = O.postprocessor_backward(gout, init, wr, orf, w9, mode, scale)
    gt = O.generator_tail_backward(g["grad_weight"], g["grad_offset"], feature, wr, cww, cow)
    for k, v in gt.items():
        ref = z[f"{prec}_{k}"]
        sk = max(1.0, float(np.abs(ref).max()))
        loose = 1 if prec == "f64" and k != "grad_feature" else 50   # f64 grad_feature is stored as float32
        _close(v, ref, max(rtol, 1e-6 if k == "grad_feature" else rtol), atol * sk * loose, k)


@pytest.mark.parametrize("path", PP, ids=[os.path.basename(p)[:-4] for p in PP])
@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_postprocessor_matches_reference(path, prec):
    z = np.load(path)
    dt = np.float64 if prec == "f64" else np.float32
    # fp64: algebraic identity with the reference; fp32: rounding-order noise only
    rtol, atol = (1e-12, 1e-12) if prec == "f64" else (1e-5, 2e-6)
    init, weight, offset, gout = (z["in_" + k].astype(dt) for k in ("init", "weight", "offset", "grad_out"))
    w9, b1 = z["in_w"].astype(dt).reshape(9), z["in_b"].astype(dt)[0]
    mode, scale = int(z["norm_mode"]), float(z["scale"])
    out = O.postprocessor_forward(init, weight, offset, w9, b1, mode, scale)
    _close(out, z[prec + "_out"], rtol, atol, "out")
    g = O.postprocessor_backward(gout, init, weight, offset, w9, mode, scale)
    for k in ("grad_init", "grad_weight", "grad_offset", "grad_w", "grad_b"):
        ref = z[f"{prec}_{k}"]
        scale_k = float(np.abs(ref).max())   # the tensor's own scale (the pp_small_grad fixtures hold gradients ~1e-6)
        _close(g[k], ref, rtol, atol * scale_k * (50 if k in ("grad_w", "grad_b") and prec == "f32" else 1), k)


@pytest.mark.parametrize("path", NL, ids=[os.path.basename(p)[:-4] for p in NL])
@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_nlspn_forward_matches_reference(path, prec):
    z = np.load(path)
    dt = np.float64 if prec == "f64" else np.float32
    rtol, atol = (1e-12, 1e-12) if prec == "f64" else (1e-5, 2e-6)
    conv_out = z[prec + "_conv_out"].astype(dt)
    conf = z["in_confidence"].astype(dt)
    gamma = z["in_gamma"].astype(dt)[0]
    offset, aff = O.nlspn_offset_affinity(conv_out, conf, gamma, affinity=str(z["cfg_affinity"]),
                                          conf_prop=bool(z["cfg_conf_prop"]), legacy=bool(z["cfg_legacy"]))
    _close(offset, z[prec + "_offset"], rtol, atol, "offset")
    _close(aff, z[prec + "_aff"], rtol, atol, "aff")
    feat, feats = O.nlspn_propagate(z["in_feat_init"].astype(dt), offset, aff, int(z["cfg_T"]),
                                    feat_fix=z["in_feat_fix"].astype(dt),
                                    preserve_input=bool(z["cfg_preserve_input"]))
    _close(np.stack(feats), z[prec + "_list_feat"], rtol, atol, "list_feat")
    _close(feat, z[prec + "_feat"], rtol, atol, "feat")


def test_known_answers():
    """Analytic cases from SURVEY.md §8c."""
    rng = np.random.default_rng(0)
    B, H, W = 2, 7, 9
    init = rng.random((B, 1, H, W))
    zero_off = np.zeros((B, 18, H, W))
    ones9 = np.ones(9)
    # (i) offsets 0, constant weight, residual => m == 0 => out = b + scale*init
    wconst = np.full((B, 9, H, W), 0.3)
    out = O.postprocessor_forward(init, wconst, zero_off, ones9, 0.25, O.NORM_RESIDUAL, 0.5)
    np.testing.assert_allclose(out, 0.25 + 0.5 * init, atol=1e-14)
    # (ii) offsets 0, sum mode, w == 1 => zero-padded 3x3 box mean
    out = O.postprocessor_forward(init, wconst, zero_off, ones9, 0.0, O.NORM_SUM)
    pad = np.pad(init[:, 0], ((0, 0), (1, 1), (1, 1)))
    box = sum(pad[:, i:i + H, j:j + W] for i in range(3) for j in range(3)) / 9
    np.testing.assert_allclose(out[:, 0], box, atol=1e-14)
    # (iii) integer offsets => shifted copies: only tap 4 active, offset (+2,-1)
    off = zero_off.copy()
    off[:, 8], off[:, 9] = 2.0, -1.0
    m = np.zeros((B, 9, H, W)); m[:, 4] = 1
    out = O.deform_gather(init, off, m, ones9, 0.0)
    exp = np.zeros((B, H, W)); exp[:, :H - 2, 1:] = init[:, 0, 2:, :W - 1]
    np.testing.assert_allclose(out[:, 0], exp, atol=1e-14)
    # (iv) sample at h in (H-1,H): only the top row of the pair is valid; h >= H => 0
    off = zero_off.copy(); off[:, 8] = 0.25
    out = O.deform_gather(init, off, m, ones9, 0.0)
    np.testing.assert_allclose(out[:, 0, H - 1], 0.75 * init[:, 0, H - 1], atol=1e-14)
    off[:, 8] = 1.0
    out = O.deform_gather(init, off, m, ones9, 0.0)
    np.testing.assert_allclose(out[:, 0, H - 1], 0.0, atol=0)
    # (v) NLSPN with zero conv output => offsets 0, aff 0, centre 1 => identity for all T
    conv_out = np.zeros((B, 24, H, W))
    offset, aff = O.nlspn_offset_affinity(conv_out, rng.random((B, 1, H, W)), 4.0)
    assert (offset == 0).all() and (aff[:, 4] == 1).all() and (np.delete(aff, 4, 1) == 0).all()
    feat, feats = O.nlspn_propagate(init, offset, aff, 6)
    for f in feats:
        np.testing.assert_array_equal(f, init)


def test_backward_is_gradient_of_forward():
    """Central finite differences in fp64 on the restatement itself."""
    rng = np.random.default_rng(1)
    B, H, W = 1, 5, 6
    init = rng.random((B, 1, H, W)); weight = rng.random((B, 9, H, W)) + 0.1
    offset = rng.normal(0, 1.5, (B, 18, H, W)); gout = rng.normal(size=(B, 1, H, W))
    w9 = 1 + 0.1 * rng.normal(size=9); b1 = 0.1
    for mode in (O.NORM_NONE, O.NORM_RESIDUAL, O.NORM_SUM):
        g = O.postprocessor_backward(gout, init, weight, offset, w9, mode, 0.7)
        f = lambda i, wt, o, w_: (O.postprocessor_forward(i, wt, o, w_, b1, mode, 0.7) * gout).sum()
        eps = 1e-6
        for name, arr, grad in (("init", init, g["grad_init"]), ("weight", weight, g["grad_weight"]),
                                ("offset", offset, g["grad_offset"])):
            idx = [tuple(rng.integers(0, s) for s in arr.shape) for _ in range(12)]
            for ix in idx:
                ap, am = arr.copy(), arr.copy()
                ap[ix] += eps; am[ix] -= eps
                args_p = dict(init=init, weight=weight, offset=offset); args_m = dict(args_p)
                args_p[name], args_m[name] = ap, am
                num = (f(args_p["init"], args_p["weight"], args_p["offset"], w9)
                       - f(args_m["init"], args_m["weight"], args_m["offset"], w9)) / (2 * eps)
                assert abs(num - grad[ix]) < 1e-6 * max(1, abs(num)), (mode, name, ix, num, grad[ix])


@pytest.mark.parametrize("path", PP, ids=[os.path.basename(p)[:-4] for p in PP])
def test_ref_port_matches_fixtures(path):
    """oracle/ref_port.py (what bench.py times as the CPU baseline) against the real reference's outputs."""
    import torch
    from oracle import ref_port
    z = np.load(path)
    t = lambda k: torch.from_numpy(z["in_" + k])
    mode, scale = int(z["norm_mode"]), float(z["scale"])
    out, gw, go, gw9, gb = ref_port.postprocessor_step(t("init"), t("weight"), t("offset"), t("w"), t("b"),
                                                       t("grad_out"), residual=(mode == 1), scale=scale)
    for got, key in ((out, "out"), (gw, "grad_weight"), (go, "grad_offset"), (gw9, "grad_w"), (gb, "grad_b")):
        np.testing.assert_allclose(got.numpy(), z["f32_" + key], rtol=1e-6, atol=1e-6, err_msg=key)


MODEL = sorted(glob.glob(os.path.join(GOLDEN, "model_*.npz")))


@pytest.mark.parametrize("path", MODEL, ids=[os.path.basename(p)[6:-4] for p in MODEL])
def test_oracle_matches_the_reference_model_run(path):
    """Model level (tests/golden/make_golden_model.py): the tensors captured inside the reference's own
    models.JSPSR.Model at the propagation boundary (models/JSPSR.py:371-375), its loss gradient, RMSE and MAE.
    The reference ran in fp32 only, so the gate is 1e-5 of each tensor's own scale; the fp32 C oracle forms tap
    positions with the reference's operation order (same cells)."""
    from oracle import c_oracle as C
    from oracle import epilogue_oracle as E
    C.build()
    z = np.load(path)
    assert str(z["meta"]).startswith("torch ") and "torchvision" in str(z["meta"])
    mode, scale = (1 if bool(z["residual"]) else 2), float(z["scale"])
    w9, b1 = z["in_w"].reshape(9), z["in_b"]
    out = C.forward(z["in_dem"], z["in_weight"], z["in_offset"], w9, b1, mode, scale)
    assert np.abs(out - z["ref_out"]).max() <= 1e-5 * np.abs(z["ref_out"]).max()
    g = C.backward(z["ref_grad_out"], z["in_dem"], z["in_weight"], z["in_offset"], w9, mode, scale, need_grad_init=False)
    gmax = np.abs(z["ref_grad_out"]).max()
    for k in ("grad_weight", "grad_offset", "grad_w", "grad_b"):
        ref = z["ref_" + k].reshape(np.asarray(g[k]).shape)
        floor = 1.2e-7 * gmax * (np.sqrt(z["ref_grad_out"].size) if k in ("grad_w", "grad_b") else 1.0)
        assert np.abs(g[k] - ref).max() <= 1e-5 * np.abs(ref).max() + floor, k
    _, _, vmin, vmax, border = (float(v) for v in z["cfg"])
    m = E.dem_metrics(z["ref_out"], z["in_hr_dem"], border, vmin, vmax, True)
    for i in range(len(m["rmse"])):
        assert f"{m['rmse'][i]:.4f}" == f"{z['ref_sample_rmse'][i]:.4f}" and f"{m['mae'][i]:.4f}" == f"{z['ref_sample_mae'][i]:.4f}"
    lo = E.multi_loss(z["ref_out"].astype(np.float64), z["in_hr_dem"].astype(np.float64))
    got = np.array([float(lo[k]) for k in ("L1", "L2", "Grad", "Total")])
    assert np.all(np.abs(got - z["ref_losses"]) <= 1e-5 * np.abs(z["ref_losses"]))


def test_fixture_provenance_is_recorded_and_uniform():
    """Every fixture names the torch / torchvision build that produced it (the reference pins torchvision 0.16, this image
    has 0.26: DESIGN.md section 2), and they all come from the same one."""
    metas = set()
    for p in PP + NL + GEN + MODEL:
        z = np.load(p)
        assert "meta" in z.files, p
        metas.add(str(z["meta"]))
    assert metas == {"torch 2.11.0+cu128 torchvision 0.26.0+cu128"}, metas
    try:
        import torchvision
    except ImportError:
        return
    # when torchvision is importable (build container and GPU box), it is the version the fixtures were made with, so
    # test_ref_port_matches_fixtures really re-runs the fixtures' operator
    assert torchvision.__version__.split("+")[0] == "0.26.0"


def test_lrru_cascade_from_the_reference_model():
    """tests/golden/cascade_lrru.npz: every tensor at the propagation boundary of the reference LRRU Model's four
    stages (models/LRRU.py:447-498).  The blend is exact (products by 0 / 1); each stage's output follows from its
    captured inputs through the Post_process_deconv oracle (unit w, zero b as the model initialises them)."""
    from oracle import spn_oracle as O
    z = np.load(os.path.join(GOLDEN, "cascade_lrru.npz"))
    d = z["d_clear"]
    assert 0.2 < float((d > 0).mean()) < 0.6
    prev = d                                             # LRRU.py:397-398: lidar = d_clear = depth
    for i in range(4):
        assert np.array_equal(O.lrru_preserve_blend(prev, d), z[f"blend{i}"]), i
        w9, b1 = np.ones(9, np.float64), np.zeros(1, np.float64)
        out = O.postprocessor_forward(z[f"blend{i}"].astype(np.float64), z[f"weight{i}"].astype(np.float64),
                                      z[f"offset{i}"].astype(np.float64), w9, b1, O.NORM_RESIDUAL, 1.0)
        ref = z[f"out{i}"]
        assert float(np.abs(out - ref).max()) <= 1e-5 * max(1.0, float(np.abs(ref).max())), i
        prev = ref
    assert np.array_equal(z["final"], z["out3"])
    assert str(z["meta"]) == "torch 2.11.0+cu128 torchvision 0.26.0+cu128"


def test_loop_backward_is_gradient_of_the_loop():
    """nlspn_propagate_backward against central differences of nlspn_propagate in fp64: a loss that touches every
    step's output, offsets small enough that no tap sits on an integer position within the step."""
    rng = np.random.default_rng(7)
    B, H, W, T = 1, 6, 7, 3
    feat = rng.random((B, 1, H, W))
    aff = 0.2 * rng.normal(size=(B, 9, H, W))
    offset = rng.normal(0, 1.2, (B, 18, H, W))
    gl = rng.normal(size=(T, B, 1, H, W))

    def loss(f_, a_, o_):
        _, fs = O.nlspn_propagate(f_, o_, a_, T)
        return sum(float((fs[t] * gl[t]).sum()) for t in range(T))
    _, feats = O.nlspn_propagate(feat, offset, aff, T)
    gf, ga, go = O.nlspn_propagate_backward(gl, feat, feats, offset, aff)
    eps = 1e-6
    for name, arr, grad in (("feat", feat, gf), ("aff", aff, ga), ("offset", offset, go)):
        for _ in range(10):
            ix = tuple(rng.integers(0, s_) for s_ in arr.shape)
            ap, am = arr.copy(), arr.copy()
            ap[ix] += eps; am[ix] -= eps
            args_p = dict(feat=feat, aff=aff, offset=offset); args_m = dict(args_p)
            args_p[name], args_m[name] = ap, am
            num = (loss(args_p["feat"], args_p["aff"], args_p["offset"]) - loss(args_m["feat"], args_m["aff"], args_m["offset"])) / (2 * eps)
            assert abs(num - grad[ix]) < 2e-6 * max(1, abs(num)), (name, ix, num, grad[ix])


def test_oracle_matches_the_reference_edsr_run():
    """tests/golden/edsr_spn.npz (make_golden_edsr.py): the third call site, models/EDSR.py:121-134 - `post_layer`
    behind a Generator with 64 feature channels - captured inside the reference's own EDSR(spn=True) run with its
    loss gradient; the Generator tail (spn.py:41-52,66-73) from the captured feature as well."""
    from oracle import c_oracle as C
    C.build()
    z = np.load(os.path.join(GOLDEN, "edsr_spn.npz"))
    assert str(z["meta"]) == "torch 2.11.0+cu128 torchvision 0.26.0+cu128"
    assert bool(z["residual"]) and float(z["scale"]) == 1.0 and z["in_feature"].shape[1] == 64
    w9, b1 = z["in_w"].reshape(9), z["in_b"]
    out = C.forward(z["in_dem"], z["in_weight"], z["in_offset"], w9, b1, 1, 1.0)
    assert np.abs(out - z["ref_out"]).max() <= 1e-5 * np.abs(z["ref_out"]).max()
    g = C.backward(z["ref_grad_out"], z["in_dem"], z["in_weight"], z["in_offset"], w9, 1, 1.0, need_grad_init=False)
    gmax = np.abs(z["ref_grad_out"]).max()
    for k in ("grad_weight", "grad_offset", "grad_w", "grad_b"):
        ref = z["ref_" + k].reshape(np.asarray(g[k]).shape)
        floor = 1.2e-7 * gmax * (np.sqrt(z["ref_grad_out"].size) if k in ("grad_w", "grad_b") else 1.0)
        assert np.abs(g[k] - ref).max() <= 1e-5 * np.abs(ref).max() + floor, k
    f64 = lambda a: a.astype(np.float64)
    weight, offset = O.generator_tail(f64(z["in_feature"]), f64(z["in_conv_weight_w"]), f64(z["in_conv_weight_b"]),
                                      f64(z["in_conv_offset_w"]), f64(z["in_conv_offset_b"]))
    assert np.abs(weight - z["in_weight"]).max() <= 2e-6 and np.abs(offset - z["in_offset"]).max() <= 2e-5
    assert np.all(offset[:, 8:10] == 0)
