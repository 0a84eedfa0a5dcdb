"""Pin the C restatement (oracle/spn_oracle.c) against the reference-generated
fixtures and against the numpy oracle on a larger random case."""
import glob
import os

import numpy as np
import pytest

from oracle import c_oracle as C
from oracle import spn_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PP = sorted(glob.glob(os.path.join(GOLDEN, "pp_*.npz")) + glob.glob(os.path.join(GOLDEN, "lrru_*.npz")))


@pytest.fixture(scope="module", autouse=True)
def _built():
    C.build()


@pytest.mark.parametrize("path", PP, ids=[os.path.basename(p)[:-4] for p in PP])
@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_c_oracle_matches_reference(path, prec):
    z = np.load(path)
    dt = np.float64 if prec == "f64" else np.float32
    rtol, atol = (1e-12, 1e-12) if prec == "f64" else (1e-5, 2e-6)
    init, weight, offset, gout = (z["in_" + k].astype(dt) for k in ("init", "weight", "offset", "grad_out"))
    w9, b1 = z["in_w"].astype(dt).reshape(9), z["in_b"].astype(dt)
    mode, scale = int(z["norm_mode"]), float(z["scale"])
    out = C.forward(init, weight, offset, w9, b1, mode, scale)
    np.testing.assert_allclose(out, z[prec + "_out"], rtol=rtol, atol=atol)
    g = C.backward(gout, init, weight, offset, w9, mode, scale)
    for k in ("grad_init", "grad_weight", "grad_offset", "grad_w", "grad_b"):
        ref = z[f"{prec}_{k}"]
        a = atol * float(np.abs(ref).max()) * (50 if k in ("grad_w", "grad_b") and prec == "f32" else 1)  # true tensor scale
        np.testing.assert_allclose(g[k], ref, rtol=rtol, atol=a, err_msg=k)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_c_oracle_matches_numpy_oracle(mode):
    rng = np.random.default_rng(5 + mode)
    B, H, W = 3, 37, 53
    init = rng.random((B, 1, H, W)); weight = rng.random((B, 9, H, W)) + 0.05
    offset = rng.normal(0, 3.0, (B, 18, H, W)); gout = rng.normal(size=(B, 1, H, W))
    w9 = 1 + 0.1 * rng.normal(size=9); b1 = np.array([0.2])
    np.testing.assert_allclose(C.forward(init, weight, offset, w9, b1, mode, 0.6),
                               O.postprocessor_forward(init, weight, offset, w9, b1[0], mode, 0.6), rtol=1e-12, atol=1e-12)
    gc = C.backward(gout, init, weight, offset, w9, mode, 0.6)
    gn = O.postprocessor_backward(gout, init, weight, offset, w9, mode, 0.6)
    for k in gn:
        np.testing.assert_allclose(gc[k], gn[k], rtol=1e-11, atol=1e-11, err_msg=k)
    assert C.backward(gout, init, weight, offset, w9, mode, 0.6, need_grad_init=False)["grad_init"] is None


def test_c_oracle_nan_inf_offsets_do_not_crash():
    B, H, W = 1, 6, 7
    rng = np.random.default_rng(0)
    init = rng.random((B, 1, H, W)).astype(np.float32); weight = rng.random((B, 9, H, W)).astype(np.float32)
    offset = rng.normal(0, 1, (B, 18, H, W)).astype(np.float32)
    offset[0, 0, 0, 0] = np.inf; offset[0, 3, 1, 1] = -np.inf; offset[0, 5, 2, 2] = 1e30; offset[0, 6, 3, 3] = np.nan
    out = C.forward(init, weight, offset, np.ones(9, np.float32), np.zeros(1, np.float32), 0, 1.0)
    ref = O.postprocessor_forward(init, weight, offset, np.ones(9, np.float32), np.float32(0), 0, 1.0)
    assert np.isnan(out[0, 0, 3, 3]) and np.isnan(ref[0, 0, 3, 3])
    ok = ~np.isnan(ref)
    np.testing.assert_allclose(out[ok], ref[ok], rtol=1e-5, atol=1e-6)
