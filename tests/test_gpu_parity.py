"""GPU parity tests: the CUDA path (through the C ABI) against
  (1) the committed fixtures produced by the reference's own modules (tests/golden),
  (2) the CPU oracles (oracle/spn_oracle.py, oracle/spn_oracle.c) on seeded inputs,
  (3) size-independent properties at BASELINE.json's sizes.

Tolerance (north_star: "within 1e-5 relative in fp32"): for every tensor,
    max|cuda - ref| <= 1e-5 * max|ref|  +  floor
i.e. relative to the tensor's OWN scale (element-wise relative error is meaningless where a value passes through zero;
a `max(1, .)` clamp would turn the gate into an absolute 1e-5 for small gradients).  `floor` is an explicit absolute
allowance for references that cancel to (near) zero although O(1) operands went into them: ABS_FLOOR = 1.2e-7 (one fp32
ulp of those operands) per 1e-5 of tolerance; tensors that scale with an upstream gradient pass
floor = ABS_FLOOR * max|grad_out| instead (the small-gradient fixtures: grad_out ~ 1e-6).  The reference values are the fp64 ones for the small
fixtures and the forward output; for gradients at larger sizes they are the fp32 oracle's:
d(bilinear)/d(position) is DISCONTINUOUS where a sample crosses an integer row/column,
and whether `float(y) + offset` lands left or right of that integer is decided by fp32
rounding of the sum (about 8e-6 of all taps at y ~ 100 differ between an fp32 and an fp64
evaluation - in torchvision's own fp32 kernels just the same).  The CUDA kernels form the
position with the reference's fp32 operation order, so they agree with the fp32 oracle
cell for cell.  bf16 I/O: the
oracle is the fp32 oracle evaluated on bf16-rounded inputs, bound 2^-7 of the
tensor scale (one bf16 output rounding is 2^-9 relative, gradients chain a few).
"""
import glob
import os
import re

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import c_oracle as C  # noqa: E402
from oracle import spn_oracle as O  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PP = sorted(glob.glob(os.path.join(GOLDEN, "pp_*.npz")) + glob.glob(os.path.join(GOLDEN, "lrru_*.npz")))
NL = sorted(glob.glob(os.path.join(GOLDEN, "nlspn_*.npz")))
GEN = sorted(glob.glob(os.path.join(GOLDEN, "gen_*.npz")))
FP32_TOL = 1e-5
BF16_TOL = 2.0 ** -7


@pytest.fixture(scope="module")
def jb():
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    import jspsr_b200
    from jspsr_b200 import _lib
    _lib.lib()  # raises if the CUDA library is not built: no fallback
    return jspsr_b200


def dev(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dtype)


ABS_FLOOR = 1.2e-7   # one fp32 ulp of O(1) operands, per 1e-5 of tolerance


def grad_floor(what, gout, tol=FP32_TOL):
    """Absolute allowance for a gradient tensor driven by the upstream gradient `gout`: one fp32 ulp of the largest
    upstream value for per-pixel tensors; for the global reductions (grad_w, grad_b, grad_gamma, convolution parameters),
    whose terms cancel, the random-walk rounding of an fp32 reduction over gout.size terms (sqrt(N) ulps)."""
    g = gout.detach().double().cpu().numpy() if isinstance(gout, torch.Tensor) else np.asarray(gout, dtype=np.float64)
    gmax = float(np.abs(g).max()) if g.size else 0.0
    reduction = re.search(r"grad_(w|b|gamma|scale|conv\w*)\b|grad of ", what) is not None
    return ABS_FLOOR * (tol / FP32_TOL) * gmax * (float(np.sqrt(g.size)) if reduction else 1.0)


def assert_close(got, ref, tol, what, floor=None, gout=None):
    got = got.detach().double().cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    scale = float(np.abs(ref).max()) if ref.size else 0.0          # the tensor's own scale, no clamp
    if floor is None and gout is not None:
        floor = grad_floor(what, gout, tol)
    if floor is None:
        floor = ABS_FLOOR * (tol / FP32_TOL)
    err = float(np.abs(got - ref).max()) if ref.size else 0.0
    assert err <= tol * scale + floor, f"{what}: max abs err {err:.3e} > {tol:.1e} * scale {scale:.3e} + floor {floor:.1e}"


def make_inputs(seed, B, H, W, sigma, dtype=np.float32):
    rng = np.random.default_rng(seed)
    init = rng.random((B, 1, H, W)).astype(dtype)
    weight = (1 / (1 + np.exp(-1.5 * rng.normal(size=(B, 9, H, W))))).astype(dtype)
    offset = np.clip(sigma * rng.normal(size=(B, 18, H, W)), -4 * sigma - 1, 4 * sigma + 1).astype(dtype)
    offset[:, 8:10] = 0
    gout = rng.normal(size=(B, 1, H, W)).astype(dtype)
    w9 = (1 + 0.2 * (rng.random(9) - 0.5)).astype(dtype)
    b1 = np.array([0.1], dtype=dtype)
    return init, weight, offset, gout, w9, b1


def run_cuda(jb, init, weight, offset, gout, w9, b1, mode, scale, dtype=torch.float32, need_init=True):
    from jspsr_b200 import functional as F
    ti, tw, to = dev(init, dtype).requires_grad_(need_init), dev(weight, dtype).requires_grad_(), dev(offset, dtype).requires_grad_()
    w = dev(w9.reshape(1, 1, 3, 3)).requires_grad_()
    b = dev(b1).requires_grad_()
    out = F.propagate(ti, tw, to, w, b, mode, scale)
    out.backward(dev(gout, dtype))
    return dict(out=out, grad_init=ti.grad, grad_weight=tw.grad, grad_offset=to.grad, grad_w=w.grad, grad_b=b.grad)


# ---------------------------------------------------------------------------
# (1) fixtures generated by the reference's own modules
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("path", PP, ids=[os.path.basename(p)[:-4] for p in PP])
def test_golden_postprocessor_modules(jb, path):
    import types
    z = np.load(path)
    name = os.path.basename(path)
    mode, scale = int(z["norm_mode"]), float(z["scale"])
    if name.startswith("lrru"):
        mod = jb.Post_process_deconv(types.SimpleNamespace(kernel_size=3, dkn_residual=(mode == 1)))
    else:
        mod = jb.PostProcessor(3, mode == 1, scale)
    mod = mod.cuda()
    with torch.no_grad():
        mod.w.copy_(dev(z["in_w"]))
        mod.b.copy_(dev(z["in_b"]))
    init = dev(z["in_init"]).requires_grad_()
    weight = dev(z["in_weight"]).requires_grad_()
    offset = dev(z["in_offset"]).requires_grad_()
    out = mod(init, weight, offset)
    out.backward(dev(z["in_grad_out"]))
    got = dict(out=out, grad_init=init.grad, grad_weight=weight.grad, grad_offset=offset.grad,
               grad_w=mod.w.grad, grad_b=mod.b.grad)
    for k, v in got.items():   # grad_out is ~1e-6 in the pp_small_grad fixtures (mean-reduced loss): floors follow it
        go = None if k == "out" else z["in_grad_out"]
        assert_close(v, z["f64_" + k], FP32_TOL, f"{name}:{k} vs reference fp64", gout=go)
        assert_close(v, z["f32_" + k], 2 * FP32_TOL, f"{name}:{k} vs reference fp32", gout=go)


@pytest.mark.parametrize("path", NL, ids=[os.path.basename(p)[:-4] for p in NL])
def test_golden_nlspn_module(jb, path):
    import types
    z = np.load(path)
    name = os.path.basename(path)
    legacy = bool(z["cfg_legacy"])
    args = types.SimpleNamespace(prop_time=int(z["cfg_T"]), affinity=str(z["cfg_affinity"]), affinity_gamma=0.5,
                                 conf_prop=bool(z["cfg_conf_prop"]), preserve_input=bool(z["cfg_preserve_input"]),
                                 legacy=legacy)
    mod = jb.NLSPN(args, 8, 1, 3, 3).cuda()
    with torch.no_grad():
        mod.conv_offset_aff.weight.copy_(dev(z["in_conv_w"]))
        mod.conv_offset_aff.bias.copy_(dev(z["in_conv_b"]))
    # cuDNN may pick TF32 for the guidance conv; the conv is not the path under test
    torch.backends.cudnn.allow_tf32 = False
    guidance = dev(z["in_guidance"]).requires_grad_(not legacy)
    confidence = dev(z["in_confidence"]).requires_grad_(not legacy)
    feat_init = dev(z["in_feat_init"]).requires_grad_(not legacy)
    feat_fix = dev(z["in_feat_fix"])
    with torch.set_grad_enabled(not legacy):
        feat, list_feat, offset, aff, gamma = mod(feat_init, guidance, confidence, feat_fix, None)
    assert len(list_feat) == int(z["cfg_T"])
    assert_close(offset, z["f64_offset"], FP32_TOL, name + ":offset")
    assert_close(aff, z["f64_aff"], FP32_TOL, name + ":aff")
    assert_close(torch.stack(list_feat), z["f64_list_feat"], FP32_TOL, name + ":list_feat")
    assert_close(feat, z["f64_feat"], FP32_TOL, name + ":feat")
    assert_close(gamma, z["in_gamma"], 0, name + ":gamma")
    if legacy:
        return
    loss = (feat * dev(z["in_grad_out"])).sum() + (list_feat[0] * dev(z["in_grad_mid"])).sum()
    loss.backward()
    assert_close(feat_init.grad, z["f64_grad_feat_init"], FP32_TOL, name + ":grad_feat_init", gout=z["in_grad_out"])
    assert_close(guidance.grad, z["f64_grad_guidance"], 2 * FP32_TOL, name + ":grad_guidance", gout=z["in_grad_out"])
    if args.conf_prop:
        assert_close(confidence.grad, z["f64_grad_confidence"], FP32_TOL, name + ":grad_confidence", gout=z["in_grad_out"])
    assert_close(mod.conv_offset_aff.weight.grad, z["f64_grad_conv_w"], 2 * FP32_TOL, name + ":grad_conv_w", gout=z["in_grad_out"])
    assert_close(mod.conv_offset_aff.bias.grad, z["f64_grad_conv_b"], 2 * FP32_TOL, name + ":grad_conv_b", gout=z["in_grad_out"])
    if args.affinity == "TGASS":
        assert_close(mod.aff_scale_const.grad, z["f64_grad_gamma"], FP32_TOL, name + ":grad_gamma", gout=z["in_grad_out"])
    else:
        assert mod.aff_scale_const.grad is None


# ---------------------------------------------------------------------------
# (2) seeded inputs against the CPU oracle (fp64 values), TMA and manual tile paths
# ---------------------------------------------------------------------------
SHAPES = [
    # B, H, W, sigma            W*4 % 16 != 0 -> manual tile loader; else TMA
    (2, 128, 128, 1.5),         # config tile, TMA
    (3, 70, 150, 1.5),          # ragged, manual loader (150*4 = 600)
    (1, 334, 334, 1.5),         # the reference's 3 m image size (utils/config.py:45-46), manual
    (1, 200, 260, 2.5),         # partial tiles in both directions, TMA
    (2, 33, 36, 8.0),           # far offsets: most taps leave the staged tile
    (1, 5, 3, 1.0),             # smaller than one warp row
    (1, 1, 1, 0.7),             # single pixel
]


@pytest.mark.parametrize("B,H,W,sigma", SHAPES)
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_forward_backward_vs_oracle(jb, B, H, W, sigma, mode):
    init, weight, offset, gout, w9, b1 = make_inputs(1000 + H * W + mode, B, H, W, sigma)
    scale = 0.75
    got = run_cuda(jb, init, weight, offset, gout, w9, b1, mode, scale)
    d = lambda a: a.astype(np.float64)
    ref_out = C.forward(d(init), d(weight), d(offset), d(w9), d(b1), mode, scale)
    ref = C.backward(gout, init, weight, offset, w9, mode, scale)   # fp32 oracle: see the module docstring
    assert_close(got["out"], C.forward(init, weight, offset, w9, b1, mode, scale), FP32_TOL, "out vs fp32 oracle")
    # vs fp64: positions are formed in fp32 (float(y) + offset), i.e. quantised to ulp(max(H, W)) like the reference's
    assert_close(got["out"], ref_out, FP32_TOL * max(1.0, max(H, W) / 64.0), "out vs fp64 oracle")
    for k in ("grad_init", "grad_weight", "grad_offset", "grad_w", "grad_b"):
        assert_close(got[k], ref[k], FP32_TOL, k, gout=gout)


def test_randomized_kernel_variants_vs_oracle(jb):
    """40 seeded random cases over every compile-time variant (rows per CTA 16/8/4/2, narrow/wide staged halo,
    TMA/manual staging, the three normalisation modes, with/without grad_init) against the fp32 C oracle."""
    rng = np.random.default_rng(2026)
    saved = {k: os.environ.get(k) for k in ("JSPSR_SPN_TILE_H", "JSPSR_SPN_HALO", "JSPSR_SPN_DISABLE_TMA")}
    try:
        for case in range(40):
            B = int(rng.integers(1, 4)); H = int(rng.integers(1, 70)); W = int(rng.choice([1, 5, 32, 64, 100, 128, 132, 200, 256, 300]))
            sigma = float(rng.choice([0.5, 1.5, 3.0, 7.0])); mode = int(rng.integers(0, 3)); need_init = bool(rng.integers(0, 2))
            os.environ["JSPSR_SPN_TILE_H"] = str(rng.choice([16, 8, 4, 2]))
            os.environ["JSPSR_SPN_HALO"] = str(rng.choice(["narrow", "wide"]))
            os.environ["JSPSR_SPN_DISABLE_TMA"] = str(rng.choice(["0", "1"]))
            init, weight, offset, gout, w9, b1 = make_inputs(5000 + case, B, H, W, sigma)
            got = run_cuda(jb, init, weight, offset, gout, w9, b1, mode, 0.6, need_init=need_init)
            ref_out = C.forward(init, weight, offset, w9, b1, mode, 0.6)
            ref = C.backward(gout, init, weight, offset, w9, mode, 0.6, need_grad_init=need_init)
            tag = f"case {case}: B={B} H={H} W={W} sigma={sigma} mode={mode} gi={need_init} env={ {k: os.environ[k] for k in saved} }"
            assert_close(got["out"], ref_out, FP32_TOL, tag + " out")
            for k in ("grad_weight", "grad_offset", "grad_w", "grad_b") + (("grad_init",) if need_init else ()):
                assert_close(got[k], ref[k], FP32_TOL, tag + " " + k, gout=gout)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_grad_init_block_floating_point_tile(jb):
    """The grad_init accumulation tile is block floating point (native integer shared-memory atomics, one
    power-of-two scale per CTA derived from the CTA's own sum of |contributions|).  Check the tolerance holds
    (a) for gradients of very small and very large magnitude (mean-reduced losses, loss scaling),
    (b) when one sample's gradients are 1e6 times another's (scales are per CTA),
    (c) bit-for-bit run to run for a single-CTA problem (integer sums do not depend on atomic order),
    (d) and that a non-finite grad_out poisons at least every cell the reference makes non-finite while other
        samples stay finite."""
    init, weight, offset, gout, w9, b1 = make_inputs(77, 2, 40, 128, 1.5)
    for mag in (1e-9, 1.0, 3e4):
        g = (gout * mag).astype(np.float32)
        got = run_cuda(jb, init, weight, offset, g, w9, b1, 1, 0.8)
        ref = C.backward(g, init, weight, offset, w9, 1, 0.8)
        scale = float(np.abs(ref["grad_init"]).max())
        err = float(np.abs(got["grad_init"].double().cpu().numpy() - ref["grad_init"]).max())
        assert err <= FP32_TOL * scale, f"magnitude {mag}: err {err:.3e} vs scale {scale:.3e}"
    g = gout.copy(); g[1] *= 1e6
    got = run_cuda(jb, init, weight, offset, g, w9, b1, 2, 1.0)
    ref = C.backward(g, init, weight, offset, w9, 2, 1.0)
    for b in range(2):
        scale = float(np.abs(ref["grad_init"][b]).max())
        err = float(np.abs(got["grad_init"][b].double().cpu().numpy() - ref["grad_init"][b]).max())
        assert err <= FP32_TOL * scale, f"sample {b}: err {err:.3e} vs scale {scale:.3e}"
    one = make_inputs(78, 1, 8, 128, 1.0)   # 8 rows x 128 columns = exactly one CTA
    a = run_cuda(jb, *one, 0, 1.0)["grad_init"]
    for _ in range(3):
        assert torch.equal(a, run_cuda(jb, *one, 0, 1.0)["grad_init"])
    g = gout.copy(); g[0, 0, 20, 64] = np.nan
    got = run_cuda(jb, init, weight, offset, g, w9, b1, 1, 0.8)["grad_init"].cpu().numpy()
    ref = C.backward(g, init, weight, offset, w9, 1, 0.8)["grad_init"]
    assert np.isnan(got[~np.isfinite(ref)]).all()
    # the poisoned region is the staged box of the CTA that owns row 20 (rows 16..23 -+ at most 14/15 halo rows)
    assert np.isfinite(got[1]).all() and np.isfinite(got[0, 0, :2]).all() and np.isfinite(got[0, 0, 39:]).all()


def test_detached_init_skips_grad_init(jb):
    """JSPSR detaches the DEM (models/JSPSR.py:372): same other gradients, no grad_init."""
    args = make_inputs(7, 2, 64, 128, 1.5)
    a = run_cuda(jb, *args, 1, 1.0, need_init=True)
    b = run_cuda(jb, *args, 1, 1.0, need_init=False)
    assert b["grad_init"] is None
    for k in ("out", "grad_weight", "grad_offset"):
        assert torch.equal(a[k], b[k]), k
    for k in ("grad_w", "grad_b"):
        assert_close(a[k], b[k].double().cpu().numpy(), 1e-6, k)


def test_tma_and_manual_tile_paths_agree_bitwise(jb):
    from jspsr_b200 import functional as F
    init, weight, offset, gout, w9, b1 = make_inputs(11, 3, 96, 256, 2.0)
    t = [dev(x) for x in (init, weight, offset)]
    w, b = dev(w9.reshape(1, 1, 3, 3)), dev(b1)
    outs, grads = [], []
    for flag in ("0", "1"):
        os.environ["JSPSR_SPN_DISABLE_TMA"] = flag
        try:
            outs.append(F.spn_forward(*t, w, b, 1, 1.0))
            grads.append(F.spn_backward(dev(gout), *t, w, 1, 1.0, need_grad_init=False))
        finally:
            os.environ["JSPSR_SPN_DISABLE_TMA"] = "0"
    assert torch.equal(outs[0], outs[1])
    assert torch.equal(grads[0][1], grads[1][1]) and torch.equal(grads[0][2], grads[1][2])


def test_nonfinite_offsets_match_reference_semantics(jb):
    from jspsr_b200 import functional as F
    init, weight, offset, gout, w9, b1 = make_inputs(3, 1, 16, 32, 1.0)
    offset[0, 0, 0, 0] = np.inf
    offset[0, 3, 1, 1] = -np.inf
    offset[0, 5, 2, 2] = 1e30
    offset[0, 6, 3, 3] = np.nan
    out = F.spn_forward(dev(init), dev(weight), dev(offset), dev(w9.reshape(1, 1, 3, 3)), dev(b1), 0, 1.0)
    ref = C.forward(init, weight, offset, w9, b1, 0, 1.0)
    got = out.cpu().numpy()
    assert np.isnan(got[0, 0, 3, 3]) and np.isnan(ref[0, 0, 3, 3])
    ok = ~np.isnan(ref)
    assert np.isfinite(got[ok]).all()
    np.testing.assert_allclose(got[ok], ref[ok], rtol=1e-5, atol=1e-5)


def test_known_answers(jb):
    from jspsr_b200 import functional as F
    B, H, W = 2, 40, 132
    rng = np.random.default_rng(0)
    init = dev(rng.random((B, 1, H, W)).astype(np.float32))
    zero_off = torch.zeros(B, 18, H, W, device="cuda")
    ones_w, zero_b = torch.ones(1, 1, 3, 3, device="cuda"), torch.zeros(1, device="cuda")
    const = torch.full((B, 9, H, W), 0.3, device="cuda")
    # offsets 0 + constant affinities + residual => m == 0 => out = b + scale*init exactly
    out = F.spn_forward(init, const, zero_off, ones_w, torch.full((1,), 0.25, device="cuda"), 1, 0.5)
    assert torch.allclose(out, 0.25 + 0.5 * init, atol=1e-6)
    # sum mode, w == 1 => zero-padded 3x3 box mean
    out = F.spn_forward(init, const, zero_off, ones_w, zero_b, 2, 1.0)
    box = torch.nn.functional.avg_pool2d(init, 3, 1, 1, count_include_pad=True)
    assert torch.allclose(out, box, atol=1e-6)
    # integer offsets => exact shifted copy (only tap 4 active, offset (+2,-1))
    off = zero_off.clone(); off[:, 8] = 2.0; off[:, 9] = -1.0
    m = torch.zeros(B, 9, H, W, device="cuda"); m[:, 4] = 1
    out = F.spn_forward(init, m, off, ones_w, zero_b, 0, 1.0)
    exp = torch.zeros_like(init); exp[:, :, :H - 2, 1:] = init[:, :, 2:, :W - 1]
    assert torch.equal(out, exp)


# ---------------------------------------------------------------------------
# iterate / NLSPN pieces against the oracle
# ---------------------------------------------------------------------------
def test_iterate_matches_oracle_loop(jb):
    from jspsr_b200 import functional as F
    rng = np.random.default_rng(5)
    B, H, W, T = 2, 48, 128, 6
    init, _, offset, _, _, _ = make_inputs(5, B, H, W, 1.5)
    aff = rng.normal(0, 0.2, (B, 9, H, W)).astype(np.float32)
    feats = F.spn_iterate(dev(init), dev(aff), dev(offset), T)
    _, ref = O.nlspn_propagate(init.astype(np.float64), offset.astype(np.float64), aff.astype(np.float64), T)
    assert_close(feats, np.stack(ref), FP32_TOL, "iterate list")


# ---------------------------------------------------------------------------
# row strips (multi-GPU sharding, emulated on one device): bit-identical to the full call
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("n_strips,W", [(2, 128), (3, 150), (8, 256)])
def test_row_strips_bitwise_equal_full(jb, n_strips, W):
    from jspsr_b200 import functional as F
    from jspsr_b200.strips import strip_bounds
    B, H = 1, 96
    init, weight, offset, _, w9, b1 = make_inputs(21, B, H, W, 1.5)
    ti, tw, to = dev(init), dev(weight), dev(offset)
    w, b = dev(w9.reshape(1, 1, 3, 3)), dev(b1)
    full = F.spn_forward(ti, tw, to, w, b, 1, 1.0)
    halo = int(np.ceil(np.abs(offset[:, 0::2]).max())) + 2
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    parts = []
    for r in range(n_strips):
        r0, r1, i0, i1 = strip_bounds(H, n_strips, r, halo)
        parts.append(F.spn_forward_strip(ti[:, :, i0:i1].contiguous(), tw[:, :, r0:r1].contiguous(),
                                         to[:, :, r0:r1].contiguous(), w, b, 1, 1.0, H, r0, i0, status))
    assert int(status.item()) == 0
    assert torch.equal(torch.cat(parts, dim=2), full)
    # a halo that is too small must be reported, never silently wrong
    r0, r1, i0, i1 = strip_bounds(H, n_strips, 0, 0)
    F.spn_forward_strip(ti[:, :, i0:i1].contiguous(), tw[:, :, r0:r1].contiguous(), to[:, :, r0:r1].contiguous(),
                        w, b, 1, 1.0, H, r0, i0, status)
    assert int(status.item()) == 1


def test_offset_absmax(jb):
    from jspsr_b200 import functional as F
    _, _, offset, _, _, _ = make_inputs(9, 3, 50, 70, 2.0)
    got = F.offset_absmax(dev(offset)).cpu().numpy()
    assert got[0] == np.abs(offset[:, 0::2]).max() and got[1] == np.abs(offset[:, 1::2]).max()


# ---------------------------------------------------------------------------
# bf16 I/O (new capability; fp32 arithmetic inside)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("mode", [1, 2])
def test_bf16_io_against_fp32_oracle_on_rounded_inputs(jb, mode):
    B, H, W = 2, 64, 128
    init, weight, offset, gout, w9, b1 = make_inputs(31 + mode, B, H, W, 1.5)
    rnd = lambda a: torch.from_numpy(a).to(torch.bfloat16).to(torch.float32).numpy()
    init, weight, offset, gout = rnd(init), rnd(weight), rnd(offset), rnd(gout)
    got = run_cuda(jb, init, weight, offset, gout, w9, b1, mode, 1.0, dtype=torch.bfloat16)
    d = lambda a: a.astype(np.float64)
    ref_out = C.forward(d(init), d(weight), d(offset), d(w9), d(b1), mode, 1.0)
    ref = C.backward(d(gout), d(init), d(weight), d(offset), d(w9), mode, 1.0)
    assert got["out"].dtype == torch.bfloat16 and got["grad_weight"].dtype == torch.bfloat16
    assert_close(got["out"], ref_out, BF16_TOL, "bf16 out")
    for k in ("grad_init", "grad_weight", "grad_offset"):
        assert_close(got[k], ref[k], BF16_TOL, "bf16 " + k)
    for k in ("grad_w", "grad_b"):  # accumulated in fp32/fp64 from exact products
        assert_close(got[k], ref[k], 1e-4, "bf16 " + k)


@pytest.mark.parametrize("io", ["bf16", "mixed"])
def test_bf16_two_pixel_and_packed_kernels_on_128x128_planes(jb, io, monkeypatch):
    """128 x 128 planes with bf16 weight / offset take the two-pixels-per-thread forward (spn_forward.cu, PAIR) and the
    packed-register backward (spn_backward.cu, PACK).  Both are the one-pixel kernels' arithmetic in another order of
    issue: bit-identical to JSPSR_SPN_PAIR=0 for every normalisation mode, tile height, on-chip and out-of-tile taps
    (sigma 16 sends two thirds of the taps through the global path), and within the bf16 bound of the fp32 oracle."""
    from jspsr_b200 import functional as F
    dt = torch.bfloat16
    dti = torch.bfloat16 if io == "bf16" else torch.float32
    for B, sigma, ths in ((3, 1.5, ("16", "8", "4")), (2, 16.0, ("16", "8")), (40, 1.5, (None,))):
        init, weight, offset, gout, w9, b1 = make_inputs(71 + B, B, 128, 128, sigma)
        if B == 2:  # non-finite and absurd offsets go through the same out-of-tile path (bits compared, NaN included)
            offset[0, 0, 5, 7], offset[0, 3, 64, 64], offset[1, 17, 127, 126], offset[1, 6, 0, 1] = np.inf, np.nan, -np.inf, 3.0e38
        ti, tg = dev(init, dti), dev(gout, dti)
        tw, to = dev(weight, dt), dev(offset, dt)
        w, b = dev(w9.reshape(1, 1, 3, 3)), dev(b1)
        for th in ths:
            if th is None:
                monkeypatch.delenv("JSPSR_SPN_TILE_H", raising=False)
            else:
                monkeypatch.setenv("JSPSR_SPN_TILE_H", th)
            for mode in (0, 1, 2):
                res = {}
                for pair in ("0", "1"):
                    monkeypatch.setenv("JSPSR_SPN_PAIR", pair)
                    res[pair] = (F.spn_forward(ti, tw, to, w, b, mode, 0.7),
                                 F.spn_backward(tg, ti, tw, to, w, mode, 0.7, need_grad_init=False))
                tag = f"{io} B={B} sigma={sigma} th={th} mode={mode}"
                bits = lambda t: t.view(torch.int16 if t.dtype == torch.bfloat16 else torch.int32)
                assert torch.equal(bits(res["0"][0]), bits(res["1"][0])), "out " + tag
                assert torch.equal(bits(res["0"][1][1]), bits(res["1"][1][1])), "grad_weight " + tag
                assert torch.equal(bits(res["0"][1][2]), bits(res["1"][1][2])), "grad_offset " + tag
                if B == 2:
                    continue  # grad_w / grad_b are NaN there
                assert_close(res["1"][1][3], res["0"][1][3].double().cpu().numpy(), 1e-6, "grad_w " + tag, gout=tg)
                assert_close(res["1"][1][4], res["0"][1][4].double().cpu().numpy(), 1e-6, "grad_b " + tag, gout=tg)
        monkeypatch.delenv("JSPSR_SPN_TILE_H", raising=False)
        monkeypatch.setenv("JSPSR_SPN_PAIR", "1")
        if B == 3:  # against the oracle on the rounded inputs: fp64 for the output, fp32 for the gradients (a tap whose
                    # fp32 position rounds onto an integer row takes the other one-sided derivative: module docstring)
            rnd = lambda a, t: torch.from_numpy(a).to(t).to(torch.float32).numpy()
            ri, rg, rw, ro = rnd(init, dti), rnd(gout, dti), rnd(weight, dt), rnd(offset, dt)
            d = lambda a: a.astype(np.float64)
            for mode in (1, 2):
                out = F.spn_forward(ti, tw, to, w, b, mode, 0.7)
                g = F.spn_backward(tg, ti, tw, to, w, mode, 0.7, need_grad_init=False)
                ref_out = C.forward(d(ri), d(rw), d(ro), d(w9), d(b1), mode, 0.7)
                ref = C.backward(rg, ri, rw, ro, w9, mode, 0.7)
                assert_close(out, ref_out, BF16_TOL if io == "bf16" else FP32_TOL, f"{io} out mode {mode}")
                assert_close(g[1], ref["grad_weight"], BF16_TOL, f"{io} grad_weight mode {mode}")
                assert_close(g[2], ref["grad_offset"], BF16_TOL, f"{io} grad_offset mode {mode}")
                assert_close(g[3], ref["grad_w"].reshape(1, 1, 3, 3), 1e-4, f"{io} grad_w mode {mode}")
                assert_close(g[4], ref["grad_b"].reshape(1), 1e-4, f"{io} grad_b mode {mode}")


# ---------------------------------------------------------------------------
# (3) properties at BASELINE.json sizes
# ---------------------------------------------------------------------------
def _device_inputs(B, H, W, sigma=1.5, seed=1234):
    g = torch.Generator(device="cuda").manual_seed(seed)
    init = torch.rand(B, 1, H, W, device="cuda", generator=g)
    weight = torch.sigmoid(1.5 * torch.randn(B, 9, H, W, device="cuda", generator=g))
    offset = (sigma * torch.randn(B, 18, H, W, device="cuda", generator=g)).clamp_(-8, 8)
    offset[:, 8:10] = 0
    return init, weight, offset


@pytest.mark.parametrize("B,H,W", [(70, 128, 128), (50, 128, 128), (1024, 128, 128), (1, 4096, 4096)])
def test_full_size_properties(jb, B, H, W):
    from jspsr_b200 import functional as F
    init, weight, offset = _device_inputs(B, H, W)
    w = (1 + 0.1 * torch.randn(1, 1, 3, 3, device="cuda"))
    b = torch.full((1,), 0.1, device="cuda")
    out = F.spn_forward(init, weight, offset, w, b, 1, 1.0)
    assert torch.isfinite(out).all()
    # samples are independent: any sub-batch reproduces its slice bit for bit
    if B > 1:
        sel = torch.tensor([0, B // 2, B - 1], device="cuda")
        sub = F.spn_forward(init[sel].contiguous(), weight[sel].contiguous(), offset[sel].contiguous(), w, b, 1, 1.0)
        assert torch.equal(sub, out[sel])
    # a few whole samples / a crop-independent sample against the CPU oracle
    nchk = min(B, 2)
    if H * W <= 128 * 128:
        ref = C.forward(init[:nchk].cpu().numpy(), weight[:nchk].cpu().numpy(), offset[:nchk].cpu().numpy(),
                        w.cpu().numpy().reshape(9), b.cpu().numpy(), 1, 1.0)
        assert_close(out[:nchk], ref, FP32_TOL, "full-size sample vs oracle")
    # linearity in the DEM for the bare operator (mode NONE, b = 0)
    zero_b = torch.zeros(1, device="cuda")
    init2 = torch.rand_like(init)
    f1 = F.spn_forward(init, weight, offset, w, zero_b, 0, 1.0)
    f2 = F.spn_forward(init2, weight, offset, w, zero_b, 0, 1.0)
    f12 = F.spn_forward(init + 0.5 * init2, weight, offset, w, zero_b, 0, 1.0)
    lin_err = (f12 - (f1 + 0.5 * f2)).abs().max().item()
    assert lin_err <= 1e-5 * max(1.0, f12.abs().max().item())
    # backward: grad_b is the sum of grad_out, grad_w is consistent with a second call
    gout = torch.randn_like(out)
    gi, gw, go, gw9, gb = F.spn_backward(gout, init, weight, offset, w, 1, 1.0, need_grad_init=False)
    assert gi is None
    ref_gb = gout.double().sum().item()
    assert abs(gb.item() - ref_gb) <= 1e-5 * max(1.0, gout.double().abs().sum().item() ** 0.5 * 10)
    # residual normalisation: grad wrt affinities sums to zero over the 9 taps
    s = gw.sum(dim=1).abs().max().item()
    assert s <= 1e-4 * max(1.0, gw.abs().max().item())
    # the centre pair of the offsets still receives its gradient (it is an input like the others)
    assert torch.isfinite(go).all()


def test_backward_full_size_vs_oracle_subset(jb):
    """B = 70 (configs/jspsr_r3_img.yml:86 batch): gradients of 3 samples checked on the CPU oracle;
    grad_w / grad_b over the whole batch checked against a float64 torch reduction of per-sample oracle terms."""
    from jspsr_b200 import functional as F
    B, H, W = 70, 128, 128
    init, weight, offset = _device_inputs(B, H, W, seed=77)
    w = (1 + 0.1 * torch.randn(1, 1, 3, 3, device="cuda"))
    gout = torch.randn(B, 1, H, W, device="cuda")
    gi, gw, go, gw9, gb = F.spn_backward(gout, init, weight, offset, w, 1, 1.0, need_grad_init=True)
    n = lambda t: t.detach().cpu().numpy()
    ref = C.backward(n(gout), n(init), n(weight), n(offset), n(w).reshape(9), 1, 1.0)  # fp32 oracle
    assert_close(gi, ref["grad_init"], FP32_TOL, "grad_init", gout=gout)
    assert_close(gw, ref["grad_weight"], FP32_TOL, "grad_weight", gout=gout)
    assert_close(go, ref["grad_offset"], FP32_TOL, "grad_offset", gout=gout)
    assert_close(gw9, ref["grad_w"], FP32_TOL, "grad_w", gout=gout)
    assert_close(gb, ref["grad_b"], FP32_TOL, "grad_b", gout=gout)


def test_end_to_end_rmse_mae_gate(jb):
    """SURVEY section 8d gate: RMSE (evaluation/metrics.py:361-396 arithmetic: 5 % border crop, clamp, exp back to
    metres) and MAE of the refined DEM are identical to printed precision (4 dp) whether the propagation ran
    in the CUDA library or in the CPU oracle, on synthetic DFC30-shaped tiles."""
    from jspsr_b200 import synth
    batch = synth.dfc30_batch(6, 128, resolution=8, seed=5, device="cuda")
    _, weight, offset, _ = synth.propagation_inputs(6, 128, 128, seed=6, device="cuda")
    pp = jb.PostProcessor(3, True, 1.0).cuda()
    with torch.no_grad():
        pp.w.mul_(0.02)   # an untrained head would swamp the DEM; keep the refinement a perturbation
        out = pp(batch["lr_dem"], weight, offset)
    n = lambda t: t.detach().cpu().numpy()
    ref = C.forward(n(batch["lr_dem"]), n(weight), n(offset), n(pp.w).reshape(9), n(pp.b), 1, 1.0)
    kw = dict(border=0.05, value_min=synth.ELEV_MIN, value_max=synth.ELEV_MAX[8], elev_log=True)
    rmse_c, mae_c = O.rmse_mae(n(out), n(batch["hr_dem"]), **kw)
    rmse_r, mae_r = O.rmse_mae(ref, n(batch["hr_dem"]), **kw)
    assert f"{rmse_c:.4f}" == f"{rmse_r:.4f}" and f"{mae_c:.4f}" == f"{mae_r:.4f}", (rmse_c, rmse_r, mae_c, mae_r)
    assert 0.0 < rmse_c < 100.0


# ---------------------------------------------------------------------------
# interface behaviour
# ---------------------------------------------------------------------------
def test_errors_and_no_cpu_fallback(jb):
    from jspsr_b200 import functional as F
    pp = jb.PostProcessor()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pp(torch.rand(1, 1, 8, 8), torch.rand(1, 9, 8, 8), torch.rand(1, 18, 8, 8))
    pp = pp.cuda()
    with pytest.raises(RuntimeError):
        pp(torch.rand(1, 1, 8, 8, device="cuda"), torch.rand(1, 9, 8, 8, device="cuda"),
           torch.rand(1, 16, 8, 8, device="cuda"))
    with pytest.raises(RuntimeError):
        pp(torch.rand(1, 1, 8, 8, device="cuda"), torch.rand(1, 8, 8, 8, device="cuda"),
           torch.rand(1, 18, 8, 8, device="cuda"))
    with pytest.raises(NotImplementedError):
        jb.PostProcessor(kernel_size=5)
    with pytest.raises(RuntimeError):
        F.spn_forward(torch.rand(1, 1, 8, 8, device="cuda", dtype=torch.float16),
                      torch.rand(1, 9, 8, 8, device="cuda", dtype=torch.float16),
                      torch.rand(1, 18, 8, 8, device="cuda", dtype=torch.float16),
                      torch.ones(9, device="cuda"), torch.zeros(1, device="cuda"), 1, 1.0)


def test_autocast_mixed_dtypes_and_empty_batch(jb):
    """Under autocast the Generator's convs emit bf16 while the DEM stays fp32; torchvision's operator promotes to
    fp32 there, and so does the replacement.  An empty batch returns an empty tensor without launching."""
    pp = jb.PostProcessor().cuda()
    init, weight, offset = _device_inputs(2, 64, 128)
    ref = pp(init, weight, offset)
    out = pp(init, weight.bfloat16(), offset.bfloat16())
    assert out.dtype == torch.float32
    ref2 = pp(init, weight.bfloat16().float(), offset.bfloat16().float())
    assert torch.equal(out, ref2) and not torch.equal(out, ref)
    # the mixed kernels (JSPSR_MIXED: bf16 weight/offset read directly, fp32 DEM) are the fp32 arithmetic on the
    # bf16-rounded values: forward bit-identical, gradients = the fp32 kernel's rounded to bf16, every tile height
    gout = torch.randn_like(init)
    saved = os.environ.get("JSPSR_SPN_TILE_H")
    try:
        for th in ("16", "8", "4", "2"):
            os.environ["JSPSR_SPN_TILE_H"] = th
            wb, ob = weight.bfloat16().requires_grad_(), offset.bfloat16().requires_grad_()
            wf, of = weight.bfloat16().float().requires_grad_(), offset.bfloat16().float().requires_grad_()
            pp.zero_grad()
            o_m = pp(init, wb, ob); o_m.backward(gout)
            gw_m, gb_m = pp.w.grad.clone(), pp.b.grad.clone()
            pp.zero_grad()
            o_f = pp(init, wf, of); o_f.backward(gout)
            assert o_m.dtype == torch.float32 and torch.equal(o_m, o_f), th
            assert wb.grad.dtype == torch.bfloat16 and torch.equal(wb.grad, wf.grad.bfloat16()), th
            assert torch.equal(ob.grad, of.grad.bfloat16()), th
            assert_close(gw_m, pp.w.grad.double().cpu().numpy(), 1e-6, "grad_w " + th, gout=gout)
            assert_close(gb_m, pp.b.grad.double().cpu().numpy(), 1e-6, "grad_b " + th, gout=gout)
    finally:
        if saved is None:
            os.environ.pop("JSPSR_SPN_TILE_H", None)
        else:
            os.environ["JSPSR_SPN_TILE_H"] = saved
    # a real autocast region: a conv producing weight/offset in bf16, the DEM in fp32
    conv = torch.nn.Conv2d(4, 27, 1).cuda()
    x = torch.randn(2, 4, 64, 128, device="cuda")
    with torch.autocast("cuda", dtype=torch.bfloat16):
        wo = conv(x)
        assert wo.dtype == torch.bfloat16
        o_a = pp(init, torch.sigmoid(wo[:, :9]), wo[:, 9:])
    assert o_a.dtype == torch.float32
    o_a.sum().backward()
    assert conv.weight.grad is not None and torch.isfinite(conv.weight.grad).all()
    empty = pp(init[:0], weight[:0].requires_grad_(), offset[:0])
    assert tuple(empty.shape) == (0, 1, 64, 128)
    empty.sum().backward()


def test_state_dict_round_trip_and_optimizer_step(jb):
    pp = jb.PostProcessor(3, True, 1.0).cuda()
    sd = pp.state_dict()
    assert set(sd) == {"w", "b"} and tuple(sd["w"].shape) == (1, 1, 3, 3) and tuple(sd["b"].shape) == (1,)
    pp2 = jb.PostProcessor().cuda()
    pp2.load_state_dict({k: v + 0.1 for k, v in sd.items()})
    init, weight, offset = _device_inputs(2, 128, 128)
    opt = torch.optim.AdamW(pp2.parameters(), lr=1e-3, weight_decay=1e-6)  # configs/*.yml:71-75
    out = pp2(init.detach(), weight.requires_grad_(), offset.requires_grad_())
    out.square().mean().backward()
    assert pp2.w.grad is not None and pp2.b.grad is not None
    before = pp2.w.detach().clone()
    opt.step()
    opt.zero_grad(set_to_none=True)
    assert not torch.equal(before, pp2.w.detach())


def test_cuda_graph_capture(jb):
    from jspsr_b200 import functional as F
    init, weight, offset = _device_inputs(2, 128, 128)
    w, b = torch.ones(1, 1, 3, 3, device="cuda"), torch.zeros(1, device="cuda")
    gout = torch.randn(2, 1, 128, 128, device="cuda")
    eager_out = F.spn_forward(init, weight, offset, w, b, 1, 1.0)
    eager_g = F.spn_backward(gout, init, weight, offset, w, 1, 1.0, need_grad_init=True)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        F.spn_backward(gout, init, weight, offset, w, 1, 1.0)  # allocate this stream's workspace before capture
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        g_out = F.spn_forward(init, weight, offset, w, b, 1, 1.0)
        g_g = F.spn_backward(gout, init, weight, offset, w, 1, 1.0, need_grad_init=True)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(g_out, eager_out)
    assert torch.equal(g_g[1], eager_g[1]) and torch.equal(g_g[2], eager_g[2])
    assert_close(g_g[0], eager_g[0].double().cpu().numpy(), 1e-6, "graph grad_init", gout=gout)
    assert_close(g_g[3], eager_g[3].double().cpu().numpy(), 1e-6, "graph grad_w", gout=gout)


def test_host_buffer_entry_point(jb):
    from jspsr_b200 import functional as F
    from jspsr_b200.host import forward_host
    init, weight, offset, _, w9, b1 = make_inputs(41, 10, 128, 128, 1.5)
    out_h = forward_host(init, weight, offset, w9, b1, 1, 1.0, chunk_B=3)
    out_d = F.spn_forward(dev(init), dev(weight), dev(offset), dev(w9.reshape(1, 1, 3, 3)), dev(b1), 1, 1.0)
    assert np.array_equal(out_h, out_d.cpu().numpy())


def test_matches_torchvision_cuda_when_available(jb):
    """The GPU incumbent (torchvision's own CUDA kernels, a library) as a second opinion."""
    tv = pytest.importorskip("torchvision.ops")
    init, weight, offset = _device_inputs(4, 128, 128)
    w = (1 + 0.1 * torch.randn(1, 1, 3, 3, device="cuda"))
    b = torch.full((1,), 0.1, device="cuda")
    pp = jb.PostProcessor().cuda()
    with torch.no_grad():
        pp.w.copy_(w); pp.b.copy_(b)
    mine = pp(init, weight, offset)
    m = weight - weight.mean(1, keepdim=True)
    ref = tv.deform_conv2d(init, offset, weight=w, bias=b, stride=(1, 1), padding=(1, 1), dilation=(1, 1), mask=m) + init
    assert_close(mine, ref.double().cpu().numpy(), FP32_TOL, "vs torchvision CUDA")


# ---------------------------------------------------------------------------
# Generator tail fused into the propagation (tcgen05 contraction)
# ---------------------------------------------------------------------------
def _gen_params(z_or_rng, C=64):
    if isinstance(z_or_rng, np.random.Generator):
        rng = z_or_rng
        return ((0.15 * rng.normal(size=(9, C, 1, 1))).astype(np.float32), (0.1 * rng.normal(size=9)).astype(np.float32),
                (0.2 * rng.normal(size=(16, C, 1, 1))).astype(np.float32), (0.3 * rng.normal(size=16)).astype(np.float32))
    z = z_or_rng
    return tuple(z["in_" + k] for k in ("conv_weight_w", "conv_weight_b", "conv_offset_w", "conv_offset_b"))


def _run_gen_cuda(init, feature, params, w9, b1, mode, scale, gout=None):
    from jspsr_b200 import functional as F
    cww, cwb, cow, cob = (dev(p).requires_grad_(gout is not None) for p in params)
    tf = dev(feature).requires_grad_(gout is not None)
    w = dev(np.asarray(w9).reshape(1, 1, 3, 3)).requires_grad_(gout is not None)
    b = dev(np.asarray(b1).reshape(1)).requires_grad_(gout is not None)
    conv_w = torch.cat((cww.flatten(1), cow.flatten(1)), 0)
    conv_b = torch.cat((cwb, cob))
    if gout is None:
        out, weight, offset = F.gen_spn_forward(dev(init), tf, conv_w, conv_b, w, b, mode, scale, True)
        assert torch.equal(out, F.gen_spn_forward(dev(init), tf, conv_w, conv_b, w, b, mode, scale, False))
        return dict(out=out, weight=weight, offset=offset)
    out = F.gen_propagate(dev(init), tf, conv_w, conv_b, w, b, mode, scale)
    out.backward(dev(gout))
    return dict(out=out, grad_feature=tf.grad, grad_conv_weight_w=cww.grad, grad_conv_weight_b=cwb.grad,
                grad_conv_offset_w=cow.grad, grad_conv_offset_b=cob.grad, grad_w=w.grad, grad_b=b.grad)


@pytest.mark.parametrize("path", GEN, ids=[os.path.basename(p)[:-4] for p in GEN])
def test_golden_generator_tail(jb, path):
    """Fixtures produced by the reference's own Generator + PostProcessor (tests/golden/make_golden_generator.py)."""
    z = np.load(path)
    mode, scale = int(z["norm_mode"]), float(z["scale"])
    fwd = _run_gen_cuda(z["in_init"], z["in_feature"], _gen_params(z), z["in_w"], z["in_b"], mode, scale)
    for k in ("out", "weight", "offset"):
        assert_close(fwd[k], z["f64_" + k], FP32_TOL, f"{k} vs the reference's fp64 run")
        assert_close(fwd[k], z["f32_" + k], FP32_TOL, f"{k} vs the reference's fp32 run")
    assert (fwd["offset"][:, 8:10] == 0).all()
    got = _run_gen_cuda(z["in_init"], z["in_feature"], _gen_params(z), z["in_w"], z["in_b"], mode, scale, z["in_grad_out"])
    assert_close(got["out"], z["f64_out"], FP32_TOL, "out (autograd path)")
    for k in ("grad_feature", "grad_conv_weight_w", "grad_conv_weight_b", "grad_conv_offset_w", "grad_conv_offset_b",
              "grad_w", "grad_b"):
        ref = z["f64_" + k]
        assert_close(got[k].reshape(ref.shape), ref, 2 * FP32_TOL, k, gout=z["in_grad_out"])   # gradients chain the 1e-5 of weight/offset


@pytest.mark.parametrize("B,H,W", [(2, 128, 128), (1, 40, 200), (3, 17, 64), (1, 9, 300), (1, 1, 1)])
@pytest.mark.parametrize("mode", [1, 2])
def test_generator_tail_vs_oracle(jb, B, H, W, mode):
    """Seeded features through the fused kernel against the numpy oracle (fp64) at tile-aligned, ragged and
    sub-tile sizes, rows-per-CTA 16 and 8, TMA and manual DEM staging."""
    rng = np.random.default_rng(900 + H * W + mode)
    init = rng.random((B, 1, H, W)).astype(np.float32)
    feature = rng.normal(size=(B, 64, H, W)).astype(np.float32)
    params = _gen_params(rng)
    w9 = (1 + 0.2 * (rng.random(9) - 0.5)).astype(np.float32); b1 = np.array([0.1], np.float32)
    d = lambda a: np.asarray(a, dtype=np.float64)
    weight, offset = O.generator_tail(d(feature), *[d(p) for p in params])
    ref = O.postprocessor_forward(d(init), weight, offset, d(w9), 0.1, mode, 0.7)
    saved = os.environ.get("JSPSR_SPN_TILE_H")
    try:
        for th in ("16", "8"):
            os.environ["JSPSR_SPN_TILE_H"] = th
            got = _run_gen_cuda(init, feature, params, w9, b1, mode, 0.7)
            assert_close(got["weight"], weight, FP32_TOL, f"weight TH={th}")
            assert_close(got["offset"], offset, FP32_TOL, f"offset TH={th}")
            assert_close(got["out"], ref, FP32_TOL * max(1.0, max(H, W) / 64.0), f"out TH={th}")
            # the gather itself is the same code as the unfused kernel: feeding it the fused kernel's own
            # weight/offset must reproduce `out` bit for bit
            from jspsr_b200 import functional as F
            same = F.spn_forward(dev(init), got["weight"], got["offset"], dev(w9.reshape(1, 1, 3, 3)), dev(b1), mode, 0.7)
            assert torch.equal(same, got["out"]), f"fused vs unfused gather TH={th}"
    finally:
        if saved is None:
            os.environ.pop("JSPSR_SPN_TILE_H", None)
        else:
            os.environ["JSPSR_SPN_TILE_H"] = saved


def test_generator_postprocess_drop_in(jb):
    """jspsr_b200.generator_postprocess on a Generator-shaped module == its unfused evaluation (own 1x1 convs,
    zero-centre insert, PostProcessor), forward and parameter gradients; unsupported shapes raise."""
    import torch.nn as nn
    import jspsr_b200
    torch.manual_seed(5)

    class Gen(nn.Module):   # structural twin of models/components/spn.py:8-75 (sub-module names and shapes)
        def __init__(self, cin=8, bc=16):
            super().__init__()
            self.kernel_size, self.num, self.idx_ref = 3, 8, 4
            mk = lambda i, o, k: nn.Sequential(nn.Conv2d(i, o, k, padding=k // 2), nn.ReLU())
            self.convd1, self.convd2 = mk(1, 2 * bc, 3), mk(2 * bc, 2 * bc, 3)
            self.convf1, self.convf2 = mk(cin, 2 * bc, 3), mk(2 * bc, 2 * bc, 3)
            self.conv, self.block = mk(4 * bc, 4 * bc, 3), mk(4 * bc, 4 * bc, 3)
            self.conv_weight = nn.Sequential(nn.Conv2d(4 * bc, 9, 1), nn.Sigmoid())
            self.conv_offset = nn.Module()
            self.conv_offset.conv = nn.Sequential(nn.Conv2d(4 * bc, 16, 1))

        def tail(self, feature):
            B, _, H, W = feature.shape
            weight = self.conv_weight(feature)
            offset = self.conv_offset.conv(feature).view(B, 8, 2, H, W)
            lo = list(torch.chunk(offset, 8, dim=1))
            lo.insert(4, torch.zeros((B, 1, 2, H, W)).type_as(offset))
            return weight, torch.cat(lo, dim=1).view(B, -1, H, W)

    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        gen = Gen().cuda()
        with torch.no_grad():
            gen.conv_offset.conv[0].weight.mul_(8.0)
        pp = jspsr_b200.PostProcessor(3, True, 1.0).cuda()
        dem = torch.rand(2, 1, 24, 136, device="cuda"); ctx = torch.randn(2, 8, 24, 136, device="cuda")
        gout = torch.randn(2, 1, 24, 136, device="cuda")
        out = jspsr_b200.generator_postprocess(gen, pp, dem, ctx)
        out.backward(gout)
        fused = {n: p.grad.clone() for n, p in list(gen.named_parameters()) + list(pp.named_parameters())}
        gen.zero_grad(); pp.zero_grad()
        feature = gen.block(gen.conv(torch.cat((gen.convd2(gen.convd1(dem)), gen.convf2(gen.convf1(ctx))), 1)))
        weight, offset = gen.tail(feature)
        ref = pp(dem.detach(), weight, offset)
        ref.backward(gout)
        assert_close(out, ref.detach().double().cpu().numpy(), FP32_TOL, "out")
        for n, p in list(gen.named_parameters()) + list(pp.named_parameters()):
            assert_close(fused[n], p.grad.double().cpu().numpy(), 5 * FP32_TOL, "grad of " + n, gout=gout)
        # both reference call sites detach the DEM before the Generator (models/JSPSR.py:372; models/EDSR.py:122-123 through
        # x.clone().detach()): no gradient reaches it at all.  detach_dem=False is for callers whose DEM carries one:
        # it then flows through the Generator body AND the propagation
        dem_j = dem.clone().requires_grad_()
        jspsr_b200.generator_postprocess(gen, pp, dem_j, ctx).backward(gout)
        assert dem_j.grad is None
        gen.zero_grad(); pp.zero_grad()
        dem_e = dem.clone().requires_grad_()
        jspsr_b200.generator_postprocess(gen, pp, dem_e, ctx, detach_dem=False).backward(gout)
        dem_r = dem.clone().requires_grad_()
        feature = gen.block(gen.conv(torch.cat((gen.convd2(gen.convd1(dem_r)), gen.convf2(gen.convf1(ctx))), 1)))
        weight, offset = gen.tail(feature)
        pp(dem_r, weight, offset).backward(gout)
        assert_close(dem_e.grad, dem_r.grad.double().cpu().numpy(), 5 * FP32_TOL, "detach_dem=False: grad of dem")
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    from jspsr_b200 import functional as F
    # C = 128 (EDSR / cat_only Generator): one CTA per SM, all 512 TMEM columns
    rng = np.random.default_rng(128)
    i128 = rng.random((2, 1, 40, 200)).astype(np.float32); f128 = rng.normal(size=(2, 128, 40, 200)).astype(np.float32)
    p128 = _gen_params(rng, C=128)
    got = _run_gen_cuda(i128, f128, p128, np.ones(9, np.float32), np.zeros(1, np.float32), 1, 1.0)
    d = lambda a_: np.asarray(a_, dtype=np.float64)
    rw, ro = O.generator_tail(d(f128), *[d(p_) for p_ in p128])
    assert_close(got["weight"], rw, FP32_TOL, "weight C=128")
    assert_close(got["offset"], ro, 2 * FP32_TOL, "offset C=128")   # twice the summation length
    same = F.spn_forward(dev(i128), got["weight"], got["offset"], dev(np.ones((1, 1, 3, 3), np.float32)), dev(np.zeros(1, np.float32)), 1, 1.0)
    assert torch.equal(same, got["out"])
    with pytest.raises(RuntimeError, match="C = 64"):
        F.gen_spn_forward(dem, torch.randn(2, 32, 24, 136, device="cuda"), torch.randn(25, 32, device="cuda"),
                          torch.randn(25, device="cuda"), pp.w, pp.b, 1, 1.0)
    with pytest.raises(RuntimeError, match="float32"):
        F.gen_spn_forward(dem.bfloat16(), torch.randn(2, 64, 24, 136, device="cuda").bfloat16(),
                          torch.randn(25, 64, device="cuda"), torch.randn(25, device="cuda"), pp.w, pp.b, 1, 1.0)
    with pytest.raises(RuntimeError, match="CUDA"):
        F.gen_spn_forward(dem.cpu(), torch.randn(2, 64, 24, 136), torch.randn(25, 64), torch.randn(25), pp.w, pp.b, 1, 1.0)


def test_generator_tail_autocast_bf16_features(jb):
    """torch.autocast: Generator.block emits bf16.  The fused kernel then takes bf16 features (exact in tf32), rounds
    weight/offset to bf16 before the gather (what the reference's autocast run propagates) and writes them as bf16:
    `out` must be bit-identical to the propagation kernel on the written tensors, weight/offset within one bf16
    rounding of the oracle on the same features, and the autograd path must run under autocast."""
    from jspsr_b200 import functional as F
    rng = np.random.default_rng(4242)
    for (B, H, W) in ((2, 128, 128), (1, 21, 200), (1, 7, 36)):
        init = rng.random((B, 1, H, W)).astype(np.float32)
        feat = dev(rng.normal(size=(B, 64, H, W)).astype(np.float32)).bfloat16()
        params = _gen_params(rng)
        conv_w = torch.cat((dev(params[0]).flatten(1), dev(params[2]).flatten(1)), 0)
        conv_b = torch.cat((dev(params[1]), dev(params[3])))
        w, b = dev((1 + 0.2 * (rng.random(9) - 0.5)).astype(np.float32).reshape(1, 1, 3, 3)), dev(np.array([0.1], np.float32))
        out, weight, offset = F.gen_spn_forward(dev(init), feat, conv_w, conv_b, w, b, 1, 0.9, True)
        assert weight.dtype == torch.bfloat16 and offset.dtype == torch.bfloat16 and out.dtype == torch.float32
        assert torch.equal(out, F.gen_spn_forward(dev(init), feat, conv_w, conv_b, w, b, 1, 0.9, False))
        assert torch.equal(out, F.spn_forward(dev(init), weight, offset, w, b, 1, 0.9)), (B, H, W)
        d = lambda a: np.asarray(a, dtype=np.float64)
        rw, ro = O.generator_tail(feat.double().cpu().numpy(), *[d(p_) for p_ in params])
        assert_close(weight.float(), rw, 2.0 ** -8, "weight (bf16)")
        assert_close(offset.float(), ro, 2.0 ** -8, "offset (bf16)")
        assert (offset[:, 8:10] == 0).all()
    import torch.nn as nn
    import jspsr_b200
    gen = nn.Module()
    gen.kernel_size = 3
    ident = nn.Identity()
    gen.convd1 = gen.convd2 = gen.convf1 = gen.convf2 = ident
    gen.conv, gen.block = nn.Conv2d(2, 64, 3, padding=1), nn.ReLU()
    gen.conv_weight = nn.Sequential(nn.Conv2d(64, 9, 1), nn.Sigmoid())
    gen.conv_offset = nn.Module(); gen.conv_offset.conv = nn.Sequential(nn.Conv2d(64, 16, 1))
    gen = gen.cuda(); pp = jspsr_b200.PostProcessor().cuda()
    dem = torch.rand(2, 1, 32, 128, device="cuda"); ctx = torch.randn(2, 1, 32, 128, device="cuda")
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = jspsr_b200.generator_postprocess(gen, pp, dem, ctx)
    assert out.dtype == torch.float32
    out.square().mean().backward()
    for n, p_ in gen.named_parameters():
        assert p_.grad is not None and p_.grad.dtype == p_.dtype and torch.isfinite(p_.grad).all(), n


def test_generator_postprocess_accepts_the_lrru_twin(jb):
    """models/LRRU.py:202-247 BasicDepthEncoder + models/LRRU.py:250-298 Post_process_deconv through generator_postprocess:
    last body block `ref`, plain nn.Conv2d heads with a functional sigmoid, `dkn_residual`, no scale - against the unfused
    evaluation (the reference's own sequence with torch's 1x1 convolutions and our propagation), forward and all gradients."""
    import types
    import torch.nn as nn
    import jspsr_b200
    torch.manual_seed(8)

    class Twin(nn.Module):   # structural twin of BasicDepthEncoder (sub-module names and shapes, bc = 16)
        def __init__(self, cin=8, bc=16):
            super().__init__()
            self.kernel_size, self.num, self.idx_ref = 3, 8, 4
            mk = lambda i, o: nn.Sequential(nn.Conv2d(i, o, 3, padding=1), nn.ReLU())
            self.convd1, self.convd2 = mk(1, 2 * bc), mk(2 * bc, 2 * bc)
            self.convf1, self.convf2 = mk(cin, 2 * bc), mk(2 * bc, 2 * bc)
            self.conv, self.ref = mk(4 * bc, 4 * bc), mk(4 * bc, 4 * bc)
            self.conv_weight = nn.Conv2d(4 * bc, 9, 1)
            self.conv_offset = nn.Conv2d(4 * bc, 16, 1)

        def forward(self, depth, context):      # LRRU.py:226-247
            B, _, H, W = depth.shape
            feature = self.ref(self.conv(torch.cat((self.convd2(self.convd1(depth)), self.convf2(self.convf1(context))), 1)))
            weight = torch.sigmoid(self.conv_weight(feature))
            lo = list(torch.chunk(self.conv_offset(feature).view(B, 8, 2, H, W), 8, dim=1))
            lo.insert(4, torch.zeros((B, 1, 2, H, W)).type_as(weight))
            return weight, torch.cat(lo, dim=1).view(B, -1, H, W)

    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        for dkn in (True, False):
            gen = Twin().cuda()
            with torch.no_grad():
                gen.conv_offset.weight.mul_(6.0)
            pp = jspsr_b200.Post_process_deconv(types.SimpleNamespace(kernel_size=3, dkn_residual=dkn)).cuda()
            depth = torch.rand(2, 1, 40, 136, device="cuda") + 0.2
            ctx = torch.randn(2, 8, 40, 136, device="cuda")
            gout = torch.randn(2, 1, 40, 136, device="cuda")
            out = jspsr_b200.generator_postprocess(gen, pp, depth, ctx)
            out.backward(gout)
            fused = {n: p.grad.clone() for n, p in list(gen.named_parameters()) + list(pp.named_parameters())}
            gen.zero_grad(); pp.zero_grad()
            weight, offset = gen(depth, ctx)
            ref = pp(depth, weight, offset)
            ref.backward(gout)
            assert_close(out, ref.detach().double().cpu().numpy(), FP32_TOL, f"LRRU twin out (dkn_residual={dkn})")
            for n, p_ in list(gen.named_parameters()) + list(pp.named_parameters()):
                assert_close(fused[n], p_.grad.double().cpu().numpy(), 5 * FP32_TOL, f"grad of LRRU twin {n}", gout=gout)
    finally:
        torch.backends.cudnn.allow_tf32 = prev


@pytest.mark.parametrize("B,C,H,W", [(2, 64, 128, 128), (1, 64, 21, 200), (3, 128, 9, 36), (1, 64, 1, 1), (1, 128, 40, 136)])
def test_generator_tail_grad_feature_kernel(jb, B, C, H, W):
    """grad_feature = gz x conv_w on tcgen05 (3-product tf32 split) against fp64, fp32 and bf16 operands."""
    from jspsr_b200 import functional as F
    rng = np.random.default_rng(70 + C + H * W)
    gz = rng.normal(size=(B, 25, H, W)).astype(np.float32)
    cw = (0.2 * rng.normal(size=(25, C))).astype(np.float32)
    ref = np.einsum("bjhw,jc->bchw", gz.astype(np.float64), cw.astype(np.float64))
    got = F.gen_tail_grad_feature(dev(gz), dev(cw))
    assert_close(got, ref, FP32_TOL, "grad_feature fp32")
    gzb = dev(gz).bfloat16()
    refb = np.einsum("bjhw,jc->bchw", gzb.double().cpu().numpy(), cw.astype(np.float64))
    gotb = F.gen_tail_grad_feature(gzb, dev(cw))
    assert gotb.dtype == torch.bfloat16
    assert_close(gotb.float(), refb, 2.0 ** -8, "grad_feature bf16")


@pytest.mark.parametrize("B,C,H,W", [(2, 64, 128, 128), (1, 64, 21, 200), (3, 128, 9, 36), (1, 64, 1, 4), (1, 128, 40, 136),
                                     (5, 128, 128, 128), (2, 128, 7, 33), (40, 128, 128, 128)])
def test_generator_tail_grad_params_kernel(jb, B, C, H, W):
    """grad_conv_w = sum_pixels gz x feature and grad_conv_b = sum_pixels gz on tcgen05 (K = pixel index, 3-product tf32
    split, short fp32 runs folded in fp64) against fp64; tolerance 1e-5 of the tensor's scale plus the random-walk
    floor of an fp32 reduction.  H*W % 4 != 0 takes the bounds-checked loader, ragged H*W the zero-filled TMA tail."""
    from jspsr_b200 import functional as F
    rng = np.random.default_rng(170 + C + H * W + B)
    gz = rng.normal(size=(B, 25, H, W)).astype(np.float32)
    gz[:, 3] *= 1e-3                                  # rows of very different magnitude share one accumulator
    feat = (rng.normal(size=(B, C, H, W)) + 0.5).astype(np.float32)   # non-zero mean: the sums do not cancel
    ref_w = np.einsum("bjhw,bchw->jc", gz.astype(np.float64), feat.astype(np.float64))
    ref_b = gz.astype(np.float64).sum(axis=(0, 2, 3))
    for rep in range(2):                              # the second call finds the workspace zeroed by the first
        gw, gb = F.gen_tail_grad_params(dev(gz), dev(feat))
        for j in range(25):                           # per row: each row of the matrix is its own tensor scale
            assert_close(gw[j], ref_w[j], FP32_TOL, f"grad_conv_w row {j} (call {rep})", gout=gz[:, j])
        assert_close(gb, ref_b, FP32_TOL, f"grad_conv_b (call {rep})", gout=gz)
    # bf16 operands (torch.autocast): exact in tf32, one product per K-step, 64-pixel blocks; fp32 results
    gzb, fb = dev(gz).bfloat16(), dev(feat).bfloat16()
    refb_w = np.einsum("bjhw,bchw->jc", gzb.double().cpu().numpy(), fb.double().cpu().numpy())
    refb_b = gzb.double().cpu().numpy().sum(axis=(0, 2, 3))
    gwb, gbb = F.gen_tail_grad_params(gzb, fb)
    assert gwb.dtype == torch.float32 and gbb.dtype == torch.float32
    for j in range(25):
        assert_close(gwb[j], refb_w[j], FP32_TOL, f"grad_conv_w row {j} (bf16 operands)", gout=gz[:, j])
    assert_close(gbb, refb_b, FP32_TOL, "grad_conv_b (bf16 operands)", gout=gz)
    only_w, none_b = F.gen_tail_grad_params(dev(gz), dev(feat), need_b=False)
    assert none_b is None and torch.equal(only_w, gw)
    # linearity in gz (a size-independent property): the kernel of the sum is the sum of the kernels to rounding
    gz2 = rng.normal(size=gz.shape).astype(np.float32)
    a, _ = F.gen_tail_grad_params(dev(gz2), dev(feat))
    both, _ = F.gen_tail_grad_params(dev(gz) + dev(gz2), dev(feat))
    assert_close(both, (gw + a).double().cpu().numpy(), 2 * FP32_TOL, "grad_conv_w linearity", gout=gz)


def test_generator_tail_grad_params_rejects(jb):
    from jspsr_b200 import functional as F
    gz = torch.randn(1, 25, 8, 8, device="cuda")
    with pytest.raises(RuntimeError, match="C = 64 and C = 128"):
        F.gen_tail_grad_params(gz, torch.randn(1, 32, 8, 8, device="cuda"))
    with pytest.raises(RuntimeError, match="one dtype"):
        F.gen_tail_grad_params(gz.bfloat16(), torch.randn(1, 64, 8, 8, device="cuda"))
    with pytest.raises(RuntimeError, match="feature must be"):
        F.gen_tail_grad_params(gz, torch.randn(2, 64, 8, 8, device="cuda"))


def test_generator_tail_kernels_capture_into_a_cuda_graph(jb):
    """The tcgen05 kernels only enqueue on the given stream (tensor maps are encoded on the host, nothing synchronises),
    so forward and feature gradient replay from a CUDA graph with identical results."""
    from jspsr_b200 import functional as F
    torch.manual_seed(3)
    init = torch.rand(4, 1, 64, 128, device="cuda"); feat = torch.randn(4, 64, 64, 128, device="cuda")
    cw = 0.15 * torch.randn(25, 64, device="cuda"); cb = 0.1 * torch.randn(25, device="cuda")
    w = torch.ones(1, 1, 3, 3, device="cuda"); b = torch.zeros(1, device="cuda")
    gz = torch.randn(4, 25, 64, 128, device="cuda")
    ref_out = F.gen_spn_forward(init, feat, cw, cb, w, b, 1, 1.0)
    ref_gf = F.gen_tail_grad_feature(gz, cw)
    ref_gw, ref_gb = F.gen_tail_grad_params(gz, feat)
    feat_loop = torch.rand(3, 1, 128, 128, device="cuda")
    aff_loop = torch.softmax(torch.randn(3, 9, 128, 128, device="cuda"), dim=1)
    off_loop = torch.randn(3, 18, 128, 128, device="cuda")
    ref_loop = F.spn_iterate(feat_loop, aff_loop, off_loop, 3)
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):                      # the per-stream reduction workspace exists before the capture
        F.gen_tail_grad_params(gz, feat)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    os.environ["JSPSR_SPN_ITER_FUSED"] = "1"        # the cluster launch of the single-launch loop is capturable too
    try:
        with torch.cuda.graph(graph, stream=s):
            out = F.gen_spn_forward(init, feat, cw, cb, w, b, 1, 1.0)
            gf = F.gen_tail_grad_feature(gz, cw)
            gw, gb = F.gen_tail_grad_params(gz, feat)
            loop = F.spn_iterate(feat_loop, aff_loop, off_loop, 3)
    finally:
        os.environ.pop("JSPSR_SPN_ITER_FUSED", None)
    for rep in range(2):                            # a replay finds the workspace zeroed by the previous one
        out.zero_(); gf.zero_(); gw.zero_(); gb.zero_(); loop.zero_()
        graph.replay(); torch.cuda.synchronize()
        assert torch.equal(out, ref_out) and torch.equal(gf, ref_gf) and torch.equal(loop, ref_loop)
        assert torch.allclose(gw, ref_gw, rtol=1e-6, atol=1e-6) and torch.allclose(gb, ref_gb, rtol=1e-6, atol=1e-6)


def test_generator_tail_full_size_properties(jb):
    """The configs' size (128 feature channels, 128x128 tiles, a batch far larger than L2): size-independent properties
    of the fused kernel - with/without the weight/offset outputs bit-identical; `out` bit-identical to the propagation
    kernel on the written tensors; weights strictly inside (0,1), centre pair zero - plus the oracle on one sample and
    the feature-gradient kernel's linearity."""
    from jspsr_b200 import functional as F
    g = torch.Generator(device="cuda").manual_seed(99)
    B, C = 768, 128
    init = torch.rand(B, 1, 128, 128, device="cuda", generator=g)
    feat = torch.randn(B, C, 128, 128, device="cuda", generator=g)
    cw = 0.1 * torch.randn(25, C, device="cuda", generator=g); cb = 0.1 * torch.randn(25, device="cuda", generator=g)
    w = torch.ones(1, 1, 3, 3, device="cuda"); b = torch.full((1,), 0.05, device="cuda")
    out, weight, offset = F.gen_spn_forward(init, feat, cw, cb, w, b, 1, 1.0, True)
    assert torch.equal(out, F.gen_spn_forward(init, feat, cw, cb, w, b, 1, 1.0, False))
    assert torch.equal(out, F.spn_forward(init, weight, offset, w, b, 1, 1.0))
    assert float(weight.min()) > 0.0 and float(weight.max()) < 1.0 and bool((offset[:, 8:10] == 0).all())
    assert bool(torch.isfinite(out).all())
    d = lambda t: t.double().cpu().numpy()
    for s_ in (0, B - 1):
        rw, ro = O.generator_tail(d(feat[s_:s_ + 1]), d(cw[:9]), d(cb[:9]), d(cw[9:]), d(cb[9:]))
        assert_close(weight[s_:s_ + 1], rw, FP32_TOL, f"weight of sample {s_}")
        assert_close(offset[s_:s_ + 1], ro, 2 * FP32_TOL, f"offset of sample {s_}")
    gz1 = torch.randn(B, 25, 128, 128, device="cuda", generator=g)
    gz2 = torch.randn(B, 25, 128, 128, device="cuda", generator=g)
    lin = F.gen_tail_grad_feature(gz1, cw) + F.gen_tail_grad_feature(gz2, cw)
    both = F.gen_tail_grad_feature(gz1 + gz2, cw)
    assert_close(both, d(lin), 2 * FP32_TOL, "grad_feature linearity")
    ref = np.einsum("jhw,jc->chw", d(gz1[5]), d(cw))
    assert_close(F.gen_tail_grad_feature(gz1[5:6], cw)[0], ref, FP32_TOL, "grad_feature of one sample")


def test_generator_tail_tma_and_manual_paths_agree_bitwise(jb):
    """JSPSR_SPN_DISABLE_TMA=1 replaces the TMA ring / DEM box of the tcgen05 kernels by bounds-checked loads (the path
    taken when rows are not 16-byte aligned): same arithmetic, bit-identical results, fp32 and bf16 features, C = 64 / 128."""
    from jspsr_b200 import functional as F
    g = torch.Generator(device="cuda").manual_seed(17)
    for C in (64, 128):
        init = torch.rand(2, 1, 40, 136, device="cuda", generator=g)
        feat = torch.randn(2, C, 40, 136, device="cuda", generator=g)
        cw = 0.15 * torch.randn(25, C, device="cuda", generator=g); cb = 0.1 * torch.randn(25, device="cuda", generator=g)
        w = torch.ones(1, 1, 3, 3, device="cuda"); b = torch.zeros(1, device="cuda")
        gz = torch.randn(2, 25, 40, 136, device="cuda", generator=g)
        res = {}
        for flag in ("0", "1"):
            os.environ["JSPSR_SPN_DISABLE_TMA"] = flag
            try:
                res[flag] = (F.gen_spn_forward(init, feat, cw, cb, w, b, 1, 1.0, True),
                             F.gen_spn_forward(init, feat.bfloat16(), cw, cb, w, b, 2, 1.0, True),
                             F.gen_tail_grad_feature(gz, cw), F.gen_tail_grad_feature(gz.bfloat16(), cw))
            finally:
                os.environ["JSPSR_SPN_DISABLE_TMA"] = "0"
        for a_, b_ in zip(res["0"], res["1"]):
            if isinstance(a_, tuple):
                assert all(torch.equal(x, y) for x, y in zip(a_, b_)), C
            else:
                assert torch.equal(a_, b_), C


@pytest.mark.parametrize("B,H,W,T,sigma", [(3, 128, 128, 6, 1.5), (2, 100, 90, 3, 1.5), (2, 128, 128, 4, 4.0), (1, 17, 128, 2, 2.0),
                                           (20, 128, 128, 2, 1.5)])
def test_iterate_fused_single_launch_is_bit_identical(jb, monkeypatch, B, H, W, T, sigma):
    """JSPSR_SPN_ITER_FUSED=1: all T applications in one launch (16-CTA cluster per sample, iteration-invariant tap state in
    registers, feature exchanged through distributed shared memory) against the T-launch loop: equal bits, including taps
    that leave the staged tile (sigma = 4: the global path reading the previous application's output) and ragged samples."""
    from jspsr_b200 import functional as F
    g = torch.Generator(device="cuda").manual_seed(11 + H + W)
    feat = torch.rand(B, 1, H, W, device="cuda", generator=g)
    aff = torch.softmax(torch.randn(B, 9, H, W, device="cuda", generator=g), dim=1)
    off = (sigma * torch.randn(B, 18, H, W, device="cuda", generator=g)).clamp_(-12, 12)
    off[:, 8:10] = 0
    ref = F.spn_iterate(feat, aff, off, T)
    monkeypatch.setenv("JSPSR_SPN_ITER_FUSED", "1")
    got = F.spn_iterate(feat, aff, off, T)
    assert torch.equal(got, ref)
    # non-finite offsets go through the same global path
    off2 = off.clone()
    off2[0, 0, H // 2, W // 3] = float("inf")
    off2[0, 3, H // 3, W // 2] = float("nan")
    got2 = F.spn_iterate(feat, aff, off2, T)
    monkeypatch.delenv("JSPSR_SPN_ITER_FUSED")
    ref2 = F.spn_iterate(feat, aff, off2, T)
    assert torch.equal(torch.isnan(got2), torch.isnan(ref2))
    assert torch.equal(torch.nan_to_num(got2), torch.nan_to_num(ref2))


def test_nlspn_loop_dtype_mixes_and_empty_loop(jb):
    """jspsr_spn_iterate has no mixed mode.  The autocast mix (fp32 feature, bf16 affinities / offsets) is
    promoted to fp32 - what torchvision's operator does in the reference - instead of being read as fp32 (garbage, out of
    bounds); the raw call rejects it; prop_time = 0 returns the input and an empty list (nlspn.py:222-235)."""
    import types
    from jspsr_b200 import functional as F
    rng = np.random.default_rng(77)
    B, H, W = 2, 40, 136
    feat = dev(rng.random((B, 1, H, W)).astype(np.float32))
    aff = dev((0.1 * rng.random((B, 9, H, W))).astype(np.float32), torch.bfloat16)
    off = dev(np.clip(1.5 * rng.normal(size=(B, 18, H, W)), -6, 6).astype(np.float32), torch.bfloat16)
    with pytest.raises(RuntimeError, match="one dtype"):
        F.spn_iterate(feat, aff, off, 3)
    got = F.iterate(feat, aff, off, 3)
    want = F.iterate(feat, aff.float(), off.float(), 3)
    assert got.dtype == torch.float32 and torch.equal(got, want)
    ref = feat.cpu().numpy()
    a32, o32 = aff.float().cpu().numpy(), off.float().cpu().numpy()
    for t in range(3):
        ref = C.forward(ref, a32, o32, np.ones(9, np.float32), np.zeros(1, np.float32), 0, 0.0)
        assert_close(got[t], ref, FP32_TOL, f"promoted loop, step {t}")
    # gradients flow through the promotion back to the bf16 leaves
    a_l, o_l, f_l = aff.clone().requires_grad_(), off.clone().requires_grad_(), feat.clone().requires_grad_()
    F.iterate(f_l, a_l, o_l, 2)[-1].sum().backward()
    assert a_l.grad.dtype == torch.bfloat16 and o_l.grad.dtype == torch.bfloat16 and f_l.grad.dtype == torch.float32
    assert torch.isfinite(a_l.grad.float()).all() and torch.isfinite(o_l.grad.float()).all()

    # the whole module under torch.autocast(bfloat16): the guidance conv emits bf16, the loop sees the mix
    args = types.SimpleNamespace(prop_time=3, affinity="TGASS", affinity_gamma=0.5, conf_prop=True, preserve_input=False,
                                 legacy=False)
    mod = jb.NLSPN(args, 8, 1, 3, 3).cuda()
    with torch.no_grad():
        mod.conv_offset_aff.weight.copy_(dev((0.05 * rng.normal(size=tuple(mod.conv_offset_aff.weight.shape))).astype(np.float32)))
    guidance = dev(rng.normal(size=(B, 8, H, W)).astype(np.float32))
    conf = dev(rng.random((B, 1, H, W)).astype(np.float32))
    with torch.no_grad():
        f32, l32, o32_, a32_, _ = mod(feat, guidance, conf)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            f16, l16, o16, a16, _ = mod(feat, guidance, conf)
    assert f16.dtype == torch.float32 and o16.dtype == torch.bfloat16 and len(l16) == 3
    assert_close(f16, f32.double().cpu().numpy(), 2.0 ** -6, "NLSPN under autocast vs fp32")

    args0 = types.SimpleNamespace(**{**vars(args), "prop_time": 0})
    mod0 = jb.NLSPN(args0, 8, 1, 3, 3).cuda()
    f0, l0, o0, a0, g0 = mod0(feat, guidance, conf)
    assert f0 is feat and l0 == [] and tuple(o0.shape) == (B, 18, H, W) and tuple(a0.shape) == (B, 9, H, W)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
@pytest.mark.parametrize("shape", [(3, 1, 128, 128), (2, 1, 33, 77), (1, 1, 5, 3)])
def test_preserve_blend_is_the_lrru_expression_bit_for_bit(jb, dtype, shape):
    """models/LRRU.py:447-451 (and its three repeats): mask from d_clear > 0, (1 - mask) * x + mask * d_clear - one kernel
    against the six torch launches of the reference's own lines, including non-finite values on either side."""
    dt = getattr(torch, dtype)
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(shape, device="cuda", generator=g).to(dt)
    d = torch.rand(shape, device="cuda", generator=g)
    d = (d * (torch.rand(shape, device="cuda", generator=g) > 0.6)).to(dt)          # sparse valid pixels, zeros elsewhere
    flat_x, flat_d = x.view(-1), d.view(-1)
    flat_x[0], flat_d[0] = float("inf"), 1.0         # (1 - 1) * inf = NaN survives
    flat_x[1], flat_d[1] = 2.0, float("nan")         # NaN > 0 is false: 1 * 2 + 0 * NaN = NaN
    flat_x[2], flat_d[2] = -0.0, 0.0
    flat_x[3], flat_d[3] = float("nan"), 0.0
    flat_x[4], flat_d[4] = 1.5, -3.0                 # negative "depth": not valid, 0 * -3 = -0
    mask = torch.sum(d > 0.0, dim=1, keepdim=True)
    mask = (mask > 0.0).type_as(d)
    ref = (1.0 - mask) * x + mask * d
    out = jb.functional.preserve_blend(x, d)
    assert out.dtype == dt and out.shape == x.shape
    same = (out.view(torch.int16 if dt == torch.bfloat16 else torch.int32) ==
            ref.view(torch.int16 if dt == torch.bfloat16 else torch.int32)) | (torch.isnan(out) & torch.isnan(ref))
    assert bool(same.all())
    x2 = x.clone()
    assert jb.functional.preserve_blend(x2, d, out=x2) is x2                       # in place
    assert torch.equal(torch.nan_to_num(x2.float()), torch.nan_to_num(out.float()))
    with pytest.raises(RuntimeError, match=r"\[B,1,H,W\]"):
        jb.functional.preserve_blend(x.expand(-1, 2, -1, -1), d.expand(-1, 2, -1, -1))


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,W,T,sigma", [(3, 128, 128, 6, 1.5), (2, 128, 128, 1, 1.5), (2, 128, 128, 8, 3.0),
                                           (2, 37, 150, 3, 1.5), (1, 64, 256, 6, 7.0), (2, 128, 128, 9, 1.5)])
@pytest.mark.parametrize("grad_rows", ["8", "16"])
def test_iterate_backward_split_matches_the_step_by_step_path(jb, monkeypatch, B, H, W, T, sigma, grad_rows):
    """nlspn.py:222-235 backward: T light carry launches + one gradient kernel (jspsr_spn_iterate_backward) against T
    applications of the full backward with accumulation - same per-step arithmetic, sums over t associated the same way.
    Covers the TMA and the manual staging, a 128 x 128 plane (compile-time stride) and others, offsets that leave the
    narrow staged tile (global-corner path in both kernels), every step's output carrying a gradient, and T = 9, which
    the split form does not take (falls back)."""
    F = jb.functional
    monkeypatch.setenv("JSPSR_ITER_GRAD_TH", grad_rows)   # both shapes of iter_grad_kernel (16 rows x 512 threads: T <= 6)
    g = torch.Generator(device="cuda").manual_seed(100 + T)
    feat = torch.rand(B, 1, H, W, device="cuda", generator=g)
    aff = (0.25 * torch.randn(B, 9, H, W, device="cuda", generator=g))
    off = (sigma * torch.randn(B, 18, H, W, device="cuda", generator=g)).clamp_(-24, 24)
    off[:, 8:10] = 0
    gl = torch.randn(T, B, 1, H, W, device="cuda", generator=g)
    res = {}
    for mode in ("steps", "split"):
        monkeypatch.setenv("JSPSR_ITER_BWD", mode)
        fa, aa, oa = (t.clone().requires_grad_(True) for t in (feat, aff, off))
        n0 = F.launch_count()
        out = F.iterate(fa, aa, oa, T)
        n1 = F.launch_count()
        out.backward(gl)
        res[mode] = (fa.grad, aa.grad, oa.grad, F.launch_count() - n1)
        assert n1 - n0 == T
    if T <= 8:
        assert res["split"][3] == T + 1 and res["steps"][3] == T      # T carry launches + one gradient kernel
    else:
        assert res["split"][3] == res["steps"][3] == T
    for name, a, b in zip(("grad_feat", "grad_aff", "grad_offset"), res["steps"], res["split"]):
        scale = float(a.abs().max())
        assert torch.isfinite(b).all(), name
        # both paths scatter the carry through the block-floating-point tile (each contribution rounded once to
        # S * 2^-30 of its CTA, spn_backward.cu) with different CTA sums S and associations, and the rounding of step t
        # is amplified by sum_k |a_k| in every later step: a few 1e-6 of the largest gradient after six steps
        assert float((a - b).abs().max()) <= 3e-5 * scale + 1e-12, (name, float((a - b).abs().max()), scale)
    # gradient w.r.t. the feature not requested: the last carry launch is skipped
    monkeypatch.setenv("JSPSR_ITER_BWD", "split")
    aa, oa = (t.clone().requires_grad_(True) for t in (aff, off))
    n1 = F.launch_count()
    F.iterate(feat, aa, oa, T).backward(gl)
    assert F.launch_count() - n1 == T + max(T - 1, 0) + 1 if T <= 8 else True
    for a, b in ((aa.grad, res["split"][1]), (oa.grad, res["split"][2])):   # (fp32 REDs of neighbouring CTAs land in any order)
        assert float((a - b).abs().max()) <= 3e-5 * float(b.abs().max())


@pytest.mark.gpu
def test_iterate_backward_and_blend_capture_into_a_cuda_graph(jb):
    """The split loop backward (memsets + T carry launches + the gradient kernel with its shared-memory opt-in) and the
    LRRU blend enqueue on the capturing stream only and replay: the gradients that depend on no atomics are bit-identical,
    the carried ones agree to the carry's rounding."""
    F = jb.functional
    g = torch.Generator(device="cuda").manual_seed(9)
    B, H, W, T = 2, 128, 128, 4
    feat = torch.rand(B, 1, H, W, device="cuda", generator=g)
    aff = 0.25 * torch.randn(B, 9, H, W, device="cuda", generator=g)
    off = (1.5 * torch.randn(B, 18, H, W, device="cuda", generator=g)).clamp_(-8, 8)
    gl = torch.randn(T, B, 1, H, W, device="cuda", generator=g)
    d = feat * (torch.rand(B, 1, H, W, device="cuda", generator=g) > 0.5)
    out = F.spn_iterate(feat, aff, off, T)
    eager = F.spn_iterate_backward(gl, feat, out, aff, off)
    eager_blend = F.preserve_blend(feat, d)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        captured = F.spn_iterate_backward(gl, feat, out, aff, off)
        captured_blend = F.preserve_blend(feat, d)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(captured_blend, eager_blend)
    for name, a, b in zip(("grad_feat", "grad_aff", "grad_offset"), eager, captured):
        assert torch.isfinite(b).all()
        assert float((a - b).abs().max()) <= 3e-5 * float(a.abs().max()), name


@pytest.mark.gpu
def test_iterate_backward_full_size_properties(jb, monkeypatch):
    """512 tiles of 128 x 128, T = 6 - the default dispatch of a large batch (split form, 16-row gradient kernel) without
    any switch: (i) agreement with the T-application path, (ii) linearity in the incoming gradient (a factor 2 shifts the
    exponent of every product and of the scatter tile's scale, nothing else), (iii) steps whose output carries no
    gradient and sends none back contribute exact zeros."""
    F = jb.functional
    for k in ("JSPSR_ITER_BWD", "JSPSR_ITER_GRAD_TH"):
        monkeypatch.delenv(k, raising=False)
    B, H, W, T = 512, 128, 128, 6
    g = torch.Generator(device="cuda").manual_seed(77)
    feat = torch.rand(B, 1, H, W, device="cuda", generator=g)
    aff = 0.1 * torch.sigmoid(1.5 * torch.randn(B, 9, H, W, device="cuda", generator=g))
    off = (1.5 * torch.randn(B, 18, H, W, device="cuda", generator=g)).clamp_(-8, 8)
    off[:, 8:10] = 0
    gl = torch.zeros(T, B, 1, H, W, device="cuda")
    gl[-1] = torch.randn(B, 1, H, W, device="cuda", generator=g)           # a loss on the last step's output
    out = F.spn_iterate(feat, aff, off, T)
    n0 = F.launch_count()
    gf, ga, go = F.spn_iterate_backward(gl, feat, out, aff, off)
    assert F.launch_count() - n0 == T + 1
    gf2, ga2, go2 = F.spn_iterate_backward(2.0 * gl, feat, out, aff, off)
    for name, a, b in (("grad_feat", gf, gf2), ("grad_aff", ga, ga2), ("grad_offset", go, go2)):
        assert float((2.0 * a - b).abs().max()) <= 1e-6 * float(b.abs().max()), name   # (RED order of neighbouring CTAs)
    monkeypatch.setenv("JSPSR_ITER_BWD", "steps")
    fa, aa, oa = (t.clone().requires_grad_(True) for t in (feat, aff, off))
    F.iterate(fa, aa, oa, T).backward(gl)
    for name, a, b in (("grad_feat", fa.grad, gf), ("grad_aff", aa.grad, ga), ("grad_offset", oa.grad, go)):
        assert float((a - b).abs().max()) <= 3e-5 * float(a.abs().max()), name
    # (iii) only the first step's output carries a gradient: nothing flows through the later steps
    gl0 = torch.zeros_like(gl)
    gl0[0] = gl[-1]
    monkeypatch.delenv("JSPSR_ITER_BWD")
    gf0, ga0, go0 = F.spn_iterate_backward(gl0, feat, out, aff, off)
    one = F.spn_backward(gl0[0], feat, aff, off, None, 0, 0.0, need_grad_init=True, need_grad_w=False)
    assert float((ga0 - one[1]).abs().max()) <= 1e-6 * float(one[1].abs().max())
    assert float((go0 - one[2]).abs().max()) <= 1e-6 * float(one[2].abs().max())
    assert float((gf0 - one[0]).abs().max()) <= 3e-5 * float(one[0].abs().max())


@pytest.mark.gpu
def test_lrru_cascade_replays_the_reference_model(jb):
    """tests/golden/cascade_lrru.npz (the reference LRRU Model's own forward, models/LRRU.py:447-498): stage by stage from
    the captured inputs - the blend bit for bit, the propagation to 1e-5 - and as a chain that only takes d_clear and
    the four (weight, offset) pairs and has to arrive at the model's output."""
    import types
    z = np.load(os.path.join(GOLDEN, "cascade_lrru.npz"))
    F = jb.functional
    mod = jb.Post_process_deconv(types.SimpleNamespace(kernel_size=3, dkn_residual=True)).cuda()
    d = dev(z["d_clear"])
    prev_ref, prev_own = d, d
    with torch.no_grad():
        for i in range(4):
            blend = F.preserve_blend(prev_ref, d)
            assert torch.equal(blend, dev(z[f"blend{i}"])), i
            w, o = dev(z[f"weight{i}"]), dev(z[f"offset{i}"])
            out = mod(blend, w, o)
            assert_close(out, z[f"out{i}"].astype(np.float64), FP32_TOL, f"cascade stage {i}")
            prev_ref = dev(z[f"out{i}"])
            prev_own = mod(F.preserve_blend(prev_own, d), w, o)
    assert_close(prev_own, z["final"].astype(np.float64), 4 * FP32_TOL, "cascade, chained")


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,W,T,sigma", [(2, 20, 36, 4, 1.5), (1, 33, 130, 3, 5.0), (1, 128, 128, 6, 1.5)])
def test_iterate_backward_split_vs_oracle(jb, monkeypatch, B, H, W, T, sigma):
    """The split loop backward against the fp64 oracle of the loop's gradient (oracle.nlspn_propagate_backward, itself
    held to central differences on the CPU): every step's output carries a gradient.  Inputs are chosen as in the other
    gradient tests: sampling positions at least 1e-3 away from integer coordinates, where the derivative jumps."""
    from oracle import spn_oracle as O
    F = jb.functional
    monkeypatch.setenv("JSPSR_ITER_BWD", "split")
    rng = np.random.default_rng(40 + T)
    feat = rng.random((B, 1, H, W)).astype(np.float32)
    aff = (0.2 * rng.normal(size=(B, 9, H, W))).astype(np.float32)
    off = np.clip(sigma * rng.normal(size=(B, 18, H, W)), -20, 20).astype(np.float32)
    frac = off - np.floor(off)
    off = np.where((frac < 1e-3) | (frac > 1 - 1e-3), np.floor(off) + 0.5, off).astype(np.float32)
    off[:, 8:10] = 0.0
    gl = rng.normal(size=(T, B, 1, H, W)).astype(np.float32)
    f64 = lambda a: a.astype(np.float64)
    _, feats = O.nlspn_propagate(f64(feat), f64(off), f64(aff), T)
    gf, ga, go = O.nlspn_propagate_backward(f64(gl), f64(feat), feats, f64(off), f64(aff))
    out = F.spn_iterate(dev(feat), dev(aff), dev(off), T)
    assert_close(out[-1], feats[-1], FP32_TOL, "loop forward")
    d_gf, d_ga, d_go = F.spn_iterate_backward(dev(gl), dev(feat), out, dev(aff), dev(off))
    assert_close(d_gf, gf, 2 * FP32_TOL, "grad_feat", gout=gl)
    assert_close(d_ga, ga, 2 * FP32_TOL, "grad_aff", gout=gl)
    assert_close(d_go, go, 2 * FP32_TOL, "grad_offset", gout=gl)
