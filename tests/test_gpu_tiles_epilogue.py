"""GPU parity tests of the components either side of the hot path, through the C ABI:

* tile scheduler + blended merge: BIT-EXACT against the fixtures made by the reference's own TileCrop / add_padding /
  gen_weight_row / gen_weight_col / copyto_add (tests/golden/tiles_reference.npz) and against the numpy oracle on
  seeded inputs (index work and float64 arithmetic in the reference's order: the bar is equality);
* loss + metric epilogue: the four losses within 1e-5 relative of the reference's fp32 MultiLoss and of the fp64
  oracle; the gradient within 1e-5 of the tensor scale, excluding the rare pixels next to a Sobel difference whose
  sign is decided by rounding (|difference| < 1e-6: the loss is not differentiable there); RMSE/MAE within 1e-6
  relative of the reference's MeterRMSE, identical at the 4 decimals the reference prints.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import epilogue_oracle as E  # noqa: E402
from oracle import tiles_oracle as T  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def jb():
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    import jspsr_b200
    from jspsr_b200 import _lib
    _lib.lib()
    return jspsr_b200


@pytest.fixture(scope="module")
def tiles_ref():
    return np.load(os.path.join(GOLDEN, "tiles_reference.npz"))


@pytest.fixture(scope="module")
def epi_ref():
    return np.load(os.path.join(GOLDEN, "epilogue_reference.npz"))


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ------------------------------------------------------------------ tile scheduler
@pytest.mark.parametrize("tag,n_tile", [("crop70", None), ("crop56", 4), ("crop129", None)])
def test_crop_matches_reference_tilecrop(jb, tiles_ref, tag, n_tile):
    size, k, stride, n = (int(v) for v in tiles_ref[tag + "_meta"])
    raster = dev(tiles_ref[tag + "_img"].transpose(2, 0, 1))
    got = jb.tiles.crop_tiles(raster, k, n_tile)
    assert got.shape == (n, 3, k, k)
    assert np.array_equal(got.cpu().numpy(), tiles_ref[tag + "_tiles"])


@pytest.mark.parametrize("tag", ["pad50", "pad100"])
def test_mirror_padding_matches_reference(jb, tiles_ref, tag):
    img, pad = tiles_ref[tag + "_img"], int(tiles_ref[tag + "_pad"][0])
    assert jb.tiles.cal_pad(*img.shape[:2]) == pad
    got = jb.tiles.add_padding(dev(img.transpose(2, 0, 1)), pad)
    assert np.array_equal(got.cpu().numpy(), tiles_ref[tag + "_padded"].transpose(2, 0, 1))
    assert np.array_equal(jb.tiles.remove_padding(got, pad).cpu().numpy(), img.transpose(2, 0, 1))


def test_crop_with_padding_and_explicit_walk_matches_oracle(jb):
    rng = np.random.default_rng(7)
    img = rng.random((100, 100, 2), dtype=np.float32)
    want = T.crop_tiles(img, 32, None, pad=14)                 # 128 -> 5 x 5 tiles of 32, stride 24
    got = jb.tiles.crop_tiles(dev(img.transpose(2, 0, 1)), 32, pad=14)
    assert want.shape == (25, 2, 32, 32) and np.array_equal(got.cpu().numpy(), want)
    # rectangular strip, explicit walk
    strip = rng.random((40, 200, 1), dtype=np.float32)
    got = jb.tiles.crop_tiles(dev(strip.transpose(2, 0, 1)), 16, stride=12, grid=(3, 16)).cpu().numpy()
    for r in range(3):
        for c in range(16):
            assert np.array_equal(got[r * 16 + c, 0], strip[12 * r:12 * r + 16, 12 * c:12 * c + 16, 0])


# ------------------------------------------------------------------ blended merge
@pytest.mark.parametrize("tag", ["merge9", "merge4", "merge9_b0"])
def test_merge_matches_reference_bitwise(jb, tiles_ref, tag):
    full, k, n, border = tiles_ref[tag + "_meta"]
    tiles = tiles_ref[tag + "_tiles_q"].astype(np.float32) * np.float32(0.25)
    got = jb.tiles.merge_tiles(dev(tiles), float(border), int(full))
    assert got.dtype == torch.float64
    got = got.cpu().numpy()
    ref = tiles_ref[tag + "_merged"]
    assert np.array_equal(got if tag == "merge9" else got[::3, ::2], ref)
    # model-shaped input [N,1,k,k] and fp32 output (= the float64 result rounded once)
    got32 = jb.tiles.merge_tiles(dev(tiles[:, None]), float(border), int(full), dtype=torch.float32).cpu().numpy()
    assert np.array_equal(got32, got.astype(np.float32))


@pytest.mark.parametrize("k,border,full,S", [(32, 0.0, 132, 3), (40, 0.1, 100, 2), (128, 0.05, 334, 4), (24, 0.0, 42, 1),
                                             (32, 0.0, 40, 2)])   # last: stride < L / 2, the generic kernel
def test_merge_matches_oracle_bitwise(jb, k, border, full, S):
    rng = np.random.default_rng(k + S)
    b, L, out, stride, n_x, p = T.merge_geometry(k, border, full)
    tiles = (1000.0 * rng.random((S, n_x * n_x, k, k)) - 100.0).astype(np.float32)
    got = jb.tiles.merge_tiles(dev(tiles), border, full).cpu().numpy()
    for s in range(S):
        assert np.array_equal(got[s], T.merge_tiles(tiles[s], border, full)), (s, n_x, p)


def test_merge_properties_at_raster_scale(jb):
    # 40 x 40 tiles of 128 (crop 6, stride 103): a 4133 x 4133 raster per call
    k, crop, stride, n = 128, 6, 103, 40
    L = k - 2 * crop
    out = stride * (n - 1) + L
    tiles = torch.full((1, n * n, k, k), 12.5, device="cuda")
    m = jb.tiles.merge_tiles(tiles, 0.05, stride=stride, grid=(n, n))
    assert m.shape == (1, out, out) and float((m - 12.5).abs().max()) < 1e-12     # ramps sum to one
    # crop -> merge round trip: tiles cut from one raster blend back to the raster (away from rounding)
    g = torch.Generator(device="cuda").manual_seed(3)
    raster = torch.rand(1, stride * (n - 1) + k, stride * (n - 1) + k, device="cuda", generator=g)
    t = jb.tiles.crop_tiles(raster, k, stride=stride, grid=(n, n))                # [n*n,1,k,k]
    back = jb.tiles.merge_tiles(t.reshape(1, n * n, k, k), 0.05, stride=stride, grid=(n, n))
    assert float((back[0] - raster[0, crop:-crop, crop:-crop].double()).abs().max()) < 1e-7
    # linearity in the tiles
    a = torch.rand(1, 9, 128, 128, device="cuda", generator=g)
    b2 = torch.rand(1, 9, 128, 128, device="cuda", generator=g)
    lhs = jb.tiles.merge_tiles(a + b2)
    rhs = jb.tiles.merge_tiles(a) + jb.tiles.merge_tiles(b2)
    assert float((lhs - rhs).abs().max()) < 1e-6


# ------------------------------------------------------------------ loss
def kink_mask(pred, gt, eps=1e-6):
    """Pixels whose gradient depends on a Sobel difference smaller than eps (3x3 neighbourhood, border folds)."""
    ds = E.spatial_gradient(pred.astype(np.float64)) - E.spatial_gradient(gt.astype(np.float64))
    small = np.abs(ds) < eps                                   # [B,C,2,H,W]: d/dx, d/dy
    H, W = pred.shape[-2:]
    if W == 1:
        small[:, :, 0] = False                                 # replicate padding: d/dx is identically 0, no kink
    if H == 1:
        small[:, :, 1] = False
    near = small.any(axis=2)                                   # [B,C,H,W]
    d0 = np.abs(pred.astype(np.float64) - gt.astype(np.float64)) < 1e-9
    pad = np.pad(near, ((0, 0), (0, 0), (1, 1), (1, 1)), mode="edge")
    H, W = near.shape[-2:]
    out = np.zeros_like(near)
    for i in range(3):
        for j in range(3):
            out |= pad[..., i:i + H, j:j + W]
    return out | d0


def check_loss(jb, pred, gt, ref_losses=None, ref_grad=None):
    from jspsr_b200 import epilogue as EP
    losses, grad = EP.loss_l1_l2_grad(dev(pred), dev(gt))
    losses = losses.cpu().numpy().astype(np.float64)
    o = E.multi_loss(pred.astype(np.float64), gt.astype(np.float64))
    for i, kname in enumerate(("L1", "L2", "Grad", "Total")):
        assert abs(losses[i] - float(o[kname])) <= 1e-5 * abs(float(o[kname])), (kname, losses[i], float(o[kname]))
        if ref_losses is not None:
            assert abs(losses[i] - ref_losses[kname]) <= 1e-5 * abs(ref_losses[kname]), (kname, losses[i], ref_losses[kname])
    want = E.multi_loss_grad(pred.astype(np.float64), gt.astype(np.float64))
    got = grad.cpu().numpy().astype(np.float64)
    ok = ~kink_mask(pred, gt)
    assert ok.mean() > 0.99 or pred.size < 64
    scale = np.abs(want).max()
    if not ok.any():
        return
    assert np.abs(got - want)[ok].max() <= 1e-5 * scale, (np.abs(got - want)[ok].max(), scale)
    if ref_grad is not None:
        assert np.abs(got - ref_grad)[ok].max() <= 1e-5 * scale
    # losses-only call (no gradient requested) gives the same numbers
    l2, g2 = EP.loss_l1_l2_grad(dev(pred), dev(gt), want_grad=False)
    # (the gradient path may run the fused march, the loss-only path the phased kernel: the fp32 partial sums are
    #  grouped differently, so the published floats can differ in the last place)
    assert g2 is None and np.allclose(l2.cpu().numpy().astype(np.float64), losses, rtol=1e-6, atol=0)


@pytest.mark.parametrize("tag", ["loss_a", "loss_b", "loss_c"])
def test_loss_matches_reference_multiloss(jb, epi_ref, tag):
    ref = {k: float(epi_ref[f"{tag}_{k}_f32"]) for k in ("L1", "L2", "Grad", "Total")}
    check_loss(jb, epi_ref[tag + "_pred"], epi_ref[tag + "_gt"], ref, epi_ref[tag + "_grad_f32"].astype(np.float64))


@pytest.mark.parametrize("B,C,H,W", [(1, 1, 1, 1), (2, 1, 2, 3), (1, 1, 1, 200), (1, 1, 70, 1), (3, 2, 17, 129), (2, 1, 130, 257),
                                     (70, 1, 128, 128), (1, 1, 334, 334),
                                     # float4 path (W % 4 == 0) with the image ending inside a tile, in x and in y
                                     (2, 1, 40, 136), (1, 2, 64, 256), (1, 1, 33, 260), (2, 1, 31, 4), (1, 1, 97, 132)])
def test_loss_matches_oracle(jb, B, C, H, W):
    rng = np.random.default_rng(B * 1000 + H + W)
    gt = rng.random((B, C, H, W)).astype(np.float32)
    pred = (gt + 0.05 * rng.normal(size=gt.shape)).astype(np.float32)
    check_loss(jb, pred, gt)


def test_multiloss_module_and_autograd(jb, epi_ref):
    pred = dev(epi_ref["loss_a_pred"]).requires_grad_()
    gt = dev(epi_ref["loss_a_gt"])
    crit = jb.MultiLoss(**{"L1": {"loss_fn": None, "weight": 1}, "L2": {"loss_fn": None, "weight": 1},
                           "Grad": {"loss_fn": None, "weight": 0.1}})
    out = crit(pred, gt)
    assert list(out) == ["L1", "L2", "Grad", "Total"]
    for k in out:
        ref = float(epi_ref[f"loss_a_{k}_f32"])
        assert abs(float(out[k]) - ref) <= 1e-5 * ref
    (3.0 * out["Total"]).backward()
    ref = 3.0 * epi_ref["loss_a_grad_f32"].astype(np.float64)
    ok = ~kink_mask(epi_ref["loss_a_pred"], epi_ref["loss_a_gt"])
    assert np.abs(pred.grad.cpu().numpy() - ref)[ok].max() <= 1e-5 * np.abs(ref).max()
    # repeated calls reuse the (self-cleaning) workspace
    again = crit(pred.detach(), gt)
    assert float(again["Total"]) == float(out["Total"])
    # zero difference: all losses and the gradient are exactly zero
    z = crit(gt.clone().requires_grad_(), gt)
    assert float(z["Total"]) == 0.0


def test_loss_in_cuda_graph(jb):
    from jspsr_b200 import epilogue as EP
    g = torch.Generator(device="cuda").manual_seed(1)
    pred, gt = torch.rand(4, 1, 128, 128, device="cuda", generator=g), torch.rand(4, 1, 128, 128, device="cuda", generator=g)
    eager, eager_grad = EP.loss_l1_l2_grad(pred, gt)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        EP.loss_l1_l2_grad(pred, gt)      # allocates this stream's workspace outside the capture
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            losses, grad = EP.loss_l1_l2_grad(pred, gt)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    assert torch.allclose(losses, eager, rtol=1e-6, atol=0) and torch.equal(grad, eager_grad)


# ------------------------------------------------------------------ metrics
@pytest.mark.parametrize("tag", ["metric_log", "metric_lin"])
def test_metrics_match_reference_meter(jb, epi_ref, tag):
    vmin, vmax, elev_log, border = epi_ref[tag + "_meta"]
    pred, gt = dev(epi_ref[tag + "_pred"]), dev(epi_ref[tag + "_gt"])
    m = jb.epilogue.dem_metrics(pred, gt, float(border), float(vmin), float(vmax), bool(elev_log))
    ref = epi_ref[tag + "_sample_rmse"]
    assert np.allclose(m["rmse"].cpu().numpy(), ref, rtol=2e-6, atol=0)
    meter = jb.MeterRMSE("local", border=float(border), value_min=float(vmin), value_max=float(vmax), verbose=False)
    for i in range(pred.shape[0]):                       # valid_batch_size 1, as the reference's loop
        meter.update(pred[i:i + 1], gt[i:i + 1], elev_log=bool(elev_log))
    assert f"{meter.get_score():.4f}" == f"{float(epi_ref[tag + '_score']):.4f}"
    o = E.dem_metrics(epi_ref[tag + "_pred"], epi_ref[tag + "_gt"], float(border), float(vmin), float(vmax), bool(elev_log))
    assert np.allclose(m["mae"].cpu().numpy(), o["mae"], rtol=2e-6, atol=0)
    assert np.allclose(m["sum_sq"].cpu().numpy(), o["sum_sq"], rtol=4e-6, atol=0)


@pytest.mark.parametrize("B,H,W,border", [(1, 8, 8, 0.0), (5, 33, 77, 0.1), (70, 128, 128, 0.05), (1, 2048, 2048, 0.05)])
def test_metrics_match_oracle(jb, B, H, W, border):
    rng = np.random.default_rng(B + H)
    gt = rng.random((B, 1, H, W)).astype(np.float32)
    pred = (gt + 0.02 * rng.normal(size=gt.shape)).astype(np.float32)
    for elev_log in (True, False):
        m = jb.epilogue.dem_metrics(dev(pred), dev(gt), border, -80.0, 929.0, elev_log)
        o = E.dem_metrics(pred, gt, border, -80.0, 929.0, elev_log)
        assert np.allclose(m["rmse"].cpu().numpy(), o["rmse"], rtol=2e-6, atol=0)
        assert np.allclose(m["mae"].cpu().numpy(), o["mae"], rtol=2e-6, atol=0)


@pytest.mark.parametrize("H,W,border", [(40, 36, 0.07), (128, 128, 0.05), (64, 260, 0.02), (37, 52, 0.1)])
def test_metrics_ignore_whatever_the_border_holds(jb, H, W, border):
    """The float4 kernel reads whole 16-byte words and drops the border columns of the first / last word by a select:
    non-finite values in the cropped border (rows and columns) must not reach the sums, whatever the alignment of the
    window's first column."""
    rng = np.random.default_rng(H * W)
    gt = rng.random((3, 1, H, W)).astype(np.float32)
    pred = (gt + 0.02 * rng.normal(size=gt.shape)).astype(np.float32)
    bh, bw = int(H * border), int(W * border)
    assert bh > 0 and bw > 0
    dirty_p, dirty_g = pred.copy(), gt.copy()
    for a, v in ((dirty_p, np.nan), (dirty_g, np.inf)):
        a[..., :bh, :] = v
        a[..., H - bh:, :] = v
        a[..., :, :bw] = v
        a[..., :, W - bw:] = v
    for elev_log in (True, False):
        clean = jb.epilogue.dem_metrics(dev(pred), dev(gt), border, -80.0, 929.0, elev_log)
        dirty = jb.epilogue.dem_metrics(dev(dirty_p), dev(dirty_g), border, -80.0, 929.0, elev_log)
        o = E.dem_metrics(pred, gt, border, -80.0, 929.0, elev_log)
        for key in ("sum_sq", "mae", "rmse"):
            assert torch.equal(clean[key], dirty[key]), key
        assert np.allclose(clean["rmse"].cpu().numpy(), o["rmse"], rtol=2e-6, atol=0)
        assert np.allclose(clean["mae"].cpu().numpy(), o["mae"], rtol=2e-6, atol=0)


# ------------------------------------------------------------------ the rows chained as the reference chains them
def _prop_inputs(rng, n, k, sigma=1.5):
    weight = (1.0 / (1.0 + np.exp(-1.5 * rng.normal(size=(n, 9, k, k))))).astype(np.float32)
    offset = np.clip(sigma * rng.normal(size=(n, 18, k, k)), -8, 8).astype(np.float32)
    offset[:, 8:10] = 0.0
    return weight, offset


def test_tiled_inference_chain_matches_oracle_chain(jb):
    """upscale_dem's walk (utils/utils.py:1583-1654): mirror-pad -> TileCrop -> per-tile propagation -> border crop +
    blended merge -> remove padding, every stage on the GPU, against the same chain of oracles."""
    from oracle import c_oracle as C
    rng = np.random.default_rng(11)
    size, k, pad, border = 100, 32, 14, 0.0                     # 128 padded -> 5 x 5 tiles of 32, stride 24
    raster = rng.random((size, size, 1), dtype=np.float32)
    want_tiles = T.crop_tiles(raster, k, None, pad=pad)          # [25,1,32,32]
    n = want_tiles.shape[0]
    weight, offset = _prop_inputs(rng, n, k)
    w9 = (np.ones(9) + rng.uniform(-0.1, 0.1, 9)).astype(np.float32)
    b1 = np.float32(0.1)
    want_out = C.forward(want_tiles, weight, offset, w9, b1, 1, 1.0)
    want = T.merge_tiles(want_out[:, 0].astype(np.float32), border, size + 2 * pad)[pad:-pad, pad:-pad]

    post = jb.PostProcessor(3, True, 1.0).cuda()
    with torch.no_grad():
        post.w.copy_(dev(w9).view(1, 1, 3, 3))
        post.b.fill_(float(b1))
        tiles = jb.tiles.crop_tiles(dev(raster.transpose(2, 0, 1)), k, pad=pad)
        assert np.array_equal(tiles.cpu().numpy(), want_tiles)
        out = post(tiles, dev(weight), dev(offset))
        merged = jb.tiles.remove_padding(jb.tiles.merge_tiles(out, border, size + 2 * pad), pad)
    assert merged.shape == (size, size) and merged.dtype == torch.float64
    err = np.abs(merged.cpu().numpy() - want).max()
    assert err <= 1e-5 * np.abs(want).max() + 1.2e-7, err      # the tensor's own scale + one ulp of the O(1) operands
    # the merge of the GPU's own tiles is bit-exact: the only difference above is the propagation's fp32 rounding
    again = T.merge_tiles(out[:, 0].cpu().numpy(), border, size + 2 * pad)[pad:-pad, pad:-pad]
    assert np.array_equal(merged.cpu().numpy(), again)


def test_training_step_through_the_fused_loss_matches_oracle_chain(jb):
    """models/JSPSR.py:372-375 + train/train_utils.py:205-214: propagation on the detached DEM -> MultiLoss ->
    backward; the gradients reaching (weight, offset, w, b) against oracle(loss gradient) -> oracle(backward)."""
    from oracle import c_oracle as C
    rng = np.random.default_rng(5)
    B, k = 3, 128
    init = rng.random((B, 1, k, k), dtype=np.float32)
    gt = np.clip(init + 0.05 * rng.normal(size=init.shape), 0, 1).astype(np.float32)
    weight, offset = _prop_inputs(rng, B, k)
    w9 = (np.ones(9) + rng.uniform(-0.1, 0.1, 9)).astype(np.float32)
    b1 = np.float32(0.05)

    post = jb.PostProcessor(3, True, 1.0).cuda()
    with torch.no_grad():
        post.w.copy_(dev(w9).view(1, 1, 3, 3))
        post.b.fill_(float(b1))
    crit = jb.MultiLoss(L1=1.0, L2=1.0, Grad=0.1)
    tw, to = dev(weight).requires_grad_(), dev(offset).requires_grad_()
    out = post(dev(init), tw, to)
    losses = crit(out, dev(gt))
    losses["Total"].backward()

    ref_out = C.forward(init, weight, offset, w9, b1, 1, 1.0)
    assert np.abs(out.detach().cpu().numpy() - ref_out).max() <= 1e-5
    # the loss gradient is evaluated at the GPU's own prediction (sign(d) is discontinuous; see kink_mask)
    pred = out.detach().cpu().numpy()
    o = E.multi_loss(pred.astype(np.float64), gt.astype(np.float64))
    assert abs(float(losses["Total"]) - float(o["Total"])) <= 1e-5 * float(o["Total"])
    g = E.multi_loss_grad(pred.astype(np.float64), gt.astype(np.float64)).astype(np.float32)
    _, got_g = jb.epilogue.loss_l1_l2_grad(out.detach(), dev(gt))
    got_g = got_g.cpu().numpy()
    ok = ~kink_mask(pred, gt)
    assert np.abs(got_g - g)[ok].max() <= 1e-5 * np.abs(g).max()
    # drive the oracle backward with the GPU's loss gradient so that the comparison isolates the propagation backward
    ref = C.backward(got_g, init, weight, offset, w9, 1, 1.0, need_grad_init=False)
    for name, got in (("grad_weight", tw.grad), ("grad_offset", to.grad), ("grad_w", post.w.grad.view(-1)),
                      ("grad_b", post.b.grad)):
        r = np.asarray(ref[name]).reshape(got.shape)
        # relative to the tensor's own scale (these gradients come from a mean-reduced loss and are ~1/N small); the
        # floor is one fp32 ulp of what went into an element: the largest upstream gradient for the per-pixel tensors, the
        # sum of their magnitudes for the two global reductions (which cancel)
        floor = 1.2e-7 * (np.abs(got_g).sum() if name in ("grad_w", "grad_b") else np.abs(got_g).max())
        assert np.abs(got.cpu().numpy() - r).max() <= 1e-5 * np.abs(r).max() + floor, name


# ------------------------------------------------------------------ full-size properties (4096 tiles: 67 Mpix, beyond L2)
def test_loss_full_size_properties(jb):
    """At the bench size the oracle is too slow; size-independent identities instead: each term alone has a closed-form
    gradient, the Sobel term's gradient sums to zero per plane (a derivative filter annihilates constants, replicate
    border included), swapping (pred, gt) keeps the losses and negates the gradient, and the sums agree with torch's."""
    from jspsr_b200 import epilogue as EP
    B, k = 4096, 128
    g = torch.Generator(device="cuda").manual_seed(17)
    gt = torch.rand(B, 1, k, k, device="cuda", generator=g)
    pred = gt + 0.05 * torch.randn(B, 1, k, k, device="cuda", generator=g)
    n = pred.numel()
    d = pred - gt
    # L2 alone: gradient 2 d / n, loss mean(d^2)
    l, gr = EP.loss_l1_l2_grad(pred, gt, 0.0, 1.0, 0.0)
    assert torch.allclose(gr, d * (2.0 / n), rtol=2e-6, atol=0)
    assert abs(float(l[1]) - float((d.double() ** 2).mean())) <= 1e-6 * float(l[1]) and float(l[3]) == float(l[1])
    # L1 alone: gradient sign(d) / n
    l, gr = EP.loss_l1_l2_grad(pred, gt, 1.0, 0.0, 0.0)
    assert torch.equal(gr, torch.sign(d) * torch.tensor(1.0 / n, device="cuda", dtype=torch.float32))
    assert abs(float(l[0]) - float(d.double().abs().mean())) <= 1e-6 * float(l[0])
    # Sobel term alone: integer multiples of w / (16 n), summing to exactly zero in every plane
    l, gr = EP.loss_l1_l2_grad(pred, gt, 0.0, 0.0, 1.0)
    units = gr.double() * (16.0 * n)
    assert float((units - units.round()).abs().max()) < 1e-3 and float(units.abs().max()) <= 16.0
    assert float(units.round().sum(dim=(1, 2, 3)).abs().max()) == 0.0
    import torch.nn.functional as TF
    kx = torch.tensor([[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]], device="cuda", dtype=torch.float64) / 8.0
    sob = TF.conv2d(TF.pad(d[:256].double(), [1, 1, 1, 1], mode="replicate"), torch.stack([kx, kx.t()])[:, None])
    l256, _ = EP.loss_l1_l2_grad(pred[:256], gt[:256], 0.0, 0.0, 1.0, want_grad=False)
    assert abs(float(l256[2]) - float(sob.abs().mean())) <= 1e-5 * float(l256[2])
    # antisymmetry
    la, ga = EP.loss_l1_l2_grad(pred, gt)
    lb, gb = EP.loss_l1_l2_grad(gt, pred)
    assert torch.allclose(la, lb, rtol=1e-6, atol=0) and torch.equal(ga, -gb)


def test_metrics_full_size_against_torch(jb):
    from jspsr_b200 import epilogue as EP
    B, k = 4096, 128
    g = torch.Generator(device="cuda").manual_seed(23)
    gt = torch.rand(B, 1, k, k, device="cuda", generator=g)
    pred = gt + 0.02 * torch.randn(B, 1, k, k, device="cuda", generator=g)
    m = EP.dem_metrics(pred, gt, 0.05, -80.0, 929.0, True)
    c = int(k * 0.05)
    lr = float(np.log(929.0 + 80.0))
    pe = torch.exp(pred[:, :, c:-c, c:-c].clamp(0, 1) * lr) - 80.0
    ge = torch.exp(gt[:, :, c:-c, c:-c] * lr) - 80.0
    dd = (pe - ge).double()
    assert torch.allclose(m["rmse"], dd.pow(2).mean(dim=(1, 2, 3)).sqrt(), rtol=2e-6, atol=0)
    assert torch.allclose(m["mae"], dd.abs().mean(dim=(1, 2, 3)), rtol=2e-6, atol=0)


def test_tiled_apply_is_the_explicit_chain(jb):
    g = torch.Generator(device="cuda").manual_seed(4)
    size, k, pad, border = 100, 32, 14, 0.1            # crop 3 per side; 128 padded -> 5 x 5 tiles, stride 24
    raster = torch.rand(2, size, size, device="cuda", generator=g)
    fn = lambda t: (t[:, :1] * 2.0 + t[:, 1:2]).contiguous()          # any per-pixel map: blending reproduces it
    got = jb.tiles.tiled_apply(fn, raster, k, pad=pad, border=border, batch=7)
    assert got.shape == (size, size) and got.dtype == torch.float64
    want = (raster[0] * 2.0 + raster[1]).double()
    assert float((got - want).abs().max()) < 1e-6
    # the explicit chain, bit for bit
    tiles = jb.tiles.crop_tiles(raster, k, pad=pad)
    merged = jb.tiles.merge_tiles(fn(tiles), border, stride=24, grid=(5, 5))
    assert torch.equal(got, merged[pad - 3:pad - 3 + size, pad - 3:pad - 3 + size])
    # no padding: the border crop is lost at the image edge, as in the reference's merge_dem (334 -> 322)
    r2 = torch.rand(1, 334, 334, device="cuda", generator=g)
    got2 = jb.tiles.tiled_apply(lambda t: t, r2, 128, border=0.05)
    assert got2.shape == (322, 322) and float((got2 - r2[0, 6:-6, 6:-6].double()).abs().max()) < 1e-6


# ------------------------------------------------------------------ the C++ torch extension over the same C ABI
def test_torch_extension_path_equals_ctypes_path_bitwise(jb):
    """jspsr_b200/_jspsr_torch.so (csrc/torch_binding.cpp) enqueues the same kernels of the same library with the
    bookkeeping in C++; the ctypes path (functional._Propagate / epilogue._Loss) must give identical bits."""
    from jspsr_b200 import _lib, functional as F, epilogue as EP
    e = _lib.ext()
    assert e is not None, "the torch extension was not built (python -c 'import __graft_entry__ as g; g.build()')"
    g = torch.Generator(device="cuda").manual_seed(8)
    B, k = 3, 128
    for dt_w, need_init in ((torch.float32, False), (torch.float32, True), (torch.bfloat16, False)):
        init = torch.rand(B, 1, k, k, device="cuda", generator=g)
        weight = torch.sigmoid(1.5 * torch.randn(B, 9, k, k, device="cuda", generator=g)).to(dt_w)
        offset = (1.5 * torch.randn(B, 18, k, k, device="cuda", generator=g)).clamp_(-8, 8).to(dt_w)
        gt = (init + 0.05 * torch.randn(B, 1, k, k, device="cuda", generator=g)).clamp_(0, 1)
        res = []
        for use_ext in (True, False):
            post = jb.PostProcessor(3, True, 0.9).cuda()
            with torch.no_grad():
                post.w.add_(0.05)
                post.b.fill_(0.02)
            i_, w_, o_ = init.clone().requires_grad_(need_init), weight.clone().requires_grad_(), offset.clone().requires_grad_()
            n0 = F.launch_count()
            if use_ext:
                out = post(i_, w_, o_)                                   # F.propagate -> extension
                total, losses = e.multi_loss(out, gt, 1.0, 1.0, 0.1)
            else:
                out = F._Propagate.apply(i_, w_, o_, post.w, post.b, 1, 0.9)
                total, losses = EP._Loss.apply(out, gt, 1.0, 1.0, 0.1)
            (2.0 * total).backward()
            assert F.launch_count() - n0 == 3                            # forward, loss, backward
            res.append([out.detach(), losses.detach(), w_.grad, o_.grad, post.w.grad, post.b.grad] +
                       ([i_.grad] if need_init else []))
        for idx, (a, b2) in enumerate(zip(*res)):
            assert a.dtype == b2.dtype and a.shape == b2.shape
            if idx == 6:       # grad_init: CTA tiles are flushed with fp32 RED (and out-of-tile taps scatter with fp32
                assert torch.allclose(a, b2, rtol=1e-5, atol=1e-8)       # atomics), whose order across CTAs is not fixed
            elif a.numel() > 16:
                assert torch.equal(a, b2)
            else:                                                        # global sums: fp64 atomics, order-dependent
                assert torch.allclose(a, b2, rtol=1e-5, atol=1e-7)
    # errors surface as RuntimeError with the library's message; CPU tensors never reach a kernel
    with pytest.raises(RuntimeError):
        e.propagate(init.cpu(), weight.cpu(), offset.cpu(), torch.ones(1, 1, 3, 3), torch.zeros(1), 1, 1.0)


@pytest.mark.parametrize("n_y,n_x,k,stride,border", [(2, 5, 32, 24, 0.1), (4, 1, 40, 30, 0.0), (1, 6, 24, 20, 0.05), (3, 2, 16, 5, 0.0)])
def test_rectangular_walks_round_trip(jb, n_y, n_x, k, stride, border):
    """Strips and rectangular rasters (explicit stride / grid): tiles cut from one raster blend back to it, for the
    2 x 2 gather kernel (stride >= L / 2) and the generic one (last case: up to four tiles overlap along an axis)."""
    g = torch.Generator(device="cuda").manual_seed(n_y * 10 + n_x)
    H, W = stride * (n_y - 1) + k, stride * (n_x - 1) + k
    raster = torch.rand(2, H, W, device="cuda", generator=g)
    tiles = jb.tiles.crop_tiles(raster, k, stride=stride, grid=(n_y, n_x))          # [n, 2, k, k]
    assert tiles.shape == (n_y * n_x, 2, k, k)
    c = int(k * border)
    per_channel = tiles.permute(1, 0, 2, 3).contiguous()                             # [S = 2, n, k, k]
    if k - 2 * c <= 2 * stride:
        merged = jb.tiles.merge_tiles(per_channel, border, stride=stride, grid=(n_y, n_x))
        assert merged.shape == (2, H - 2 * c, W - 2 * c)
        want = raster[:, c:H - c, c:W - c].double()
        assert float((merged - want).abs().max()) < 1e-6
    else:
        # more than two tiles overlap: the reference's ramps no longer sum to one; compare the two kernels' common
        # ground instead - every pixel covered by a single tile is an exact copy
        merged = jb.tiles.merge_tiles(per_channel, border, stride=stride, grid=(n_y, n_x))
        L = k - 2 * c
        assert merged.shape == (2, stride * (n_y - 1) + L, stride * (n_x - 1) + L)
        assert torch.equal(merged[:, :stride, :stride], raster[:, c:c + stride, c:c + stride].double())


@pytest.mark.parametrize("B,C,H,W", [(1, 1, 1, 1), (2, 1, 2, 3), (1, 1, 70, 1), (3, 2, 17, 129), (2, 1, 130, 257), (6, 1, 128, 128),
                                     (1, 1, 334, 334), (2, 1, 40, 136), (1, 2, 64, 256), (1, 1, 65, 260), (1, 1, 97, 132)])
def test_loss_matches_oracle_with_64_row_tiles(jb, monkeypatch, B, C, H, W):
    """The 64-row instantiation (picked by itself only for large batches) against the oracle on the ragged shapes."""
    monkeypatch.setenv("JSPSR_LOSS_TILE_H", "64")
    rng = np.random.default_rng(B * 999 + H + W)
    gt = rng.random((B, C, H, W)).astype(np.float32)
    pred = (gt + 0.05 * rng.normal(size=gt.shape)).astype(np.float32)
    check_loss(jb, pred, gt)
    # and it is the 32-row kernel's gradient, bit for bit (the per-pixel arithmetic does not depend on the tiling)
    from jspsr_b200 import epilogue as EP
    _, g64 = EP.loss_l1_l2_grad(dev(pred), dev(gt))
    monkeypatch.setenv("JSPSR_LOSS_TILE_H", "32")
    _, g32 = EP.loss_l1_l2_grad(dev(pred), dev(gt))
    assert torch.equal(g64, g32)


@pytest.mark.parametrize("tile_h", ["32", "64"])
@pytest.mark.parametrize("B,C,H,W", [(2, 1, 1, 4), (1, 1, 2, 8), (3, 1, 31, 4), (2, 1, 40, 136), (1, 2, 64, 256), (1, 1, 65, 260),
                                     (1, 1, 97, 132), (6, 1, 128, 128), (1, 1, 33, 128), (1, 1, 200, 384)])
def test_fused_loss_march_equals_phased_kernel_bitwise(jb, monkeypatch, tile_h, B, C, H, W):
    """The float4 gradient path runs as one per-warp march (signs and adjoint rows in registers, ring applied in
    registers); the phased kernel (sign tiles in shared memory, ring pass) computes the same per-pixel arithmetic."""
    from jspsr_b200 import epilogue as EP
    monkeypatch.setenv("JSPSR_LOSS_TILE_H", tile_h)
    rng = np.random.default_rng(H * 7 + W)
    gt = rng.random((B, C, H, W)).astype(np.float32)
    pred = (gt + 0.05 * rng.normal(size=gt.shape)).astype(np.float32)
    if W >= 128 and H >= 33:
        pred[0, 0, :8, :2] = gt[0, 0, :8, :2]                   # exact zeros of d and of the Sobel difference (corner patch)
    monkeypatch.setenv("JSPSR_LOSS_FUSED", "1")
    lf, gf = EP.loss_l1_l2_grad(dev(pred), dev(gt))
    monkeypatch.setenv("JSPSR_LOSS_FUSED", "0")
    lp, gp = EP.loss_l1_l2_grad(dev(pred), dev(gt))
    assert torch.equal(gf, gp)
    assert torch.allclose(lf, lp, rtol=1e-6, atol=0)
    monkeypatch.setenv("JSPSR_LOSS_FUSED", "1")
    check_loss(jb, pred, gt)
