"""ctypes loader for oracle/spn_oracle.c (TEST INFRASTRUCTURE ONLY).

The C restatement exists so that parity at BASELINE.json's full sizes and the
``cpu_baseline`` timing do not depend on numpy's speed; it is pinned against the
same golden fixtures as the numpy oracle (tests/test_oracle_c.py).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libspn_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "spn_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libspn_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.spn_oracle_threads.restype = ctypes.c_int
    return _lib


def threads() -> int:
    return int(lib().spn_oracle_threads())


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _sfx(dt):
    return {np.dtype(np.float32): ("f32", ctypes.c_float), np.dtype(np.float64): ("f64", ctypes.c_double)}[np.dtype(dt)]


def forward(init, weight, offset, w9, b1, mode=1, scale=1.0):
    dt = init.dtype
    sfx, creal = _sfx(dt)
    B, _, H, W = init.shape
    init, weight, offset = (np.ascontiguousarray(a, dtype=dt) for a in (init, weight, offset))
    w9 = np.ascontiguousarray(np.asarray(w9, dtype=dt).reshape(9))
    b1 = np.ascontiguousarray(np.asarray(b1, dtype=dt).reshape(1))
    out = np.empty((B, 1, H, W), dtype=dt)
    fn = getattr(lib(), "spn_oracle_forward_" + sfx)
    fn(_p(init), _p(weight), _p(offset), _p(w9), _p(b1), _p(out), ctypes.c_long(B), ctypes.c_long(H),
       ctypes.c_long(W), ctypes.c_int(mode), creal(scale))
    return out


def backward(grad_out, init, weight, offset, w9, mode=1, scale=1.0, need_grad_init=True):
    dt = init.dtype
    sfx, creal = _sfx(dt)
    B, _, H, W = init.shape
    grad_out, init, weight, offset = (np.ascontiguousarray(a, dtype=dt) for a in (grad_out, init, weight, offset))
    w9 = np.ascontiguousarray(np.asarray(w9, dtype=dt).reshape(9))
    gi = np.empty((B, 1, H, W), dtype=dt) if need_grad_init else None
    gw = np.empty_like(weight)
    go = np.empty_like(offset)
    gw9 = np.empty(9, dtype=dt)
    gb = np.empty(1, dtype=dt)
    fn = getattr(lib(), "spn_oracle_backward_" + sfx)
    fn(_p(grad_out), _p(init), _p(weight), _p(offset), _p(w9), _p(gi), _p(gw), _p(go), _p(gw9), _p(gb),
       ctypes.c_long(B), ctypes.c_long(H), ctypes.c_long(W), ctypes.c_int(mode), creal(scale))
    return dict(grad_init=gi, grad_weight=gw, grad_offset=go, grad_w=gw9.reshape(1, 1, 3, 3), grad_b=gb)
