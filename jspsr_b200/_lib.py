"""ctypes binding of libjspsr_spn.so (the C ABI in include/jspsr_spn.h, include/jspsr_tiles.h and include/jspsr_peer.h).

There is no fallback: if the library is missing or a call fails, a RuntimeError
is raised with the library's own message.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import c_float, c_int, c_size_t, c_uint, c_void_p

_CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB_PATH = os.path.join(_CSRC, "libjspsr_spn.so")

NORM_NONE, NORM_RESIDUAL, NORM_SUM = 0, 1, 2
F32, BF16, MIXED = 0, 1, 2  # MIXED: bf16 weight/offset (+ gradients), fp32 init/out/grad_out (torch.autocast)
AFFINITY = {"AS": 0, "ASS": 1, "TC": 2, "TGASS": 3}
BWD_ACCUMULATE = 1
BWD_GEN_PREACT = 2

# name -> (restype, argtypes); mirrors include/jspsr_spn.h and include/jspsr_tiles.h one to one
_SIGNATURES = {
    "jspsr_version": (c_int, []),
    "jspsr_last_error": (ctypes.c_char_p, []),
    "jspsr_spn_workspace_bytes": (c_size_t, []),
    "jspsr_spn_forward": (c_int, [c_void_p] * 6 + [c_int, c_int, c_int, c_int, c_float, c_int, c_void_p]),
    "jspsr_spn_backward": (c_int, [c_void_p] * 11 + [c_int, c_int, c_int, c_int, c_float, c_int, c_uint, c_void_p]),
    "jspsr_spn_forward_strip": (c_int, [c_void_p] * 6 + [c_int] * 8 + [c_float, c_int, c_void_p, c_void_p]),
    "jspsr_gen_spn_forward": (c_int, [c_void_p] * 9 + [c_int] * 5 + [c_float, c_int, c_void_p]),
    "jspsr_gen_tail_grad_feature": (c_int, [c_void_p] * 3 + [c_int] * 5 + [c_void_p]),
    "jspsr_gen_tail_workspace_bytes": (c_size_t, []),
    "jspsr_gen_tail_grad_params": (c_int, [c_void_p] * 5 + [c_int] * 5 + [c_void_p]),
    "jspsr_spn_offset_absmax": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "jspsr_spn_iterate_backward": (c_int, [c_void_p] * 9 + [c_int] * 5 + [c_void_p]),
    "jspsr_preserve_blend": (c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_longlong, c_int, c_void_p]),
    "jspsr_spn_iterate": (c_int, [c_void_p] * 7 + [c_int] * 5 + [c_void_p]),
    "jspsr_nlspn_affinity_forward": (c_int, [c_void_p] * 5 + [c_int] * 6 + [c_void_p]),
    "jspsr_nlspn_affinity_backward": (c_int, [c_void_p] * 9 + [c_int] * 5 + [c_void_p]),
    "jspsr_spn_host_scratch_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    # include/jspsr_tiles.h
    "jspsr_tiles_crop": (c_int, [c_void_p, c_void_p] + [c_int] * 8 + [c_void_p]),
    "jspsr_tiles_merge": (c_int, [c_void_p, c_void_p] + [c_int] * 7 + [c_void_p]),
    "jspsr_loss_l1_l2_grad": (c_int, [c_void_p, c_void_p, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p,
                                      c_int, c_int, c_int, c_void_p]),
    "jspsr_dem_metrics": (c_int, [c_void_p] * 3 + [c_int] * 5 + [c_float, c_float, c_int, c_void_p]),
    # include/jspsr_peer.h
    "jspsr_peer_alloc": (c_int, [c_size_t, c_void_p, c_void_p]),
    "jspsr_peer_open": (c_int, [c_void_p, c_void_p]),
    "jspsr_peer_close": (c_int, [c_void_p]),
    "jspsr_peer_free": (c_int, [c_void_p]),
    "jspsr_spn_backward_reduce": (c_int, [c_void_p] * 11 + [c_int, c_int, c_int, c_int, c_float, c_int, c_uint, c_void_p,
                                          c_void_p]),
    "jspsr_spn_forward_strip_peer": (c_int, [c_void_p] * 6 + [c_int] * 7 + [c_float, c_int, c_void_p, c_void_p, c_void_p]),
    "jspsr_strip_halo_push": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "jspsr_spn_forward_host": (c_int, [c_void_p] * 6 + [c_int, c_int, c_int, c_int, c_float, c_int, c_void_p, c_size_t, c_int]),
}

_lib = None


EXT_PATH = os.path.join(os.path.dirname(_CSRC), "_jspsr_torch.so")
_EXT_SRC = os.path.join(_CSRC, "torch_binding.cpp")
_ext = None
_ext_tried = False


def build(verbose: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU), then the thin torch
    extension over its C ABI (csrc/torch_binding.cpp: C++ autograd wrappers, no kernels of its own)."""
    cmd = ["make", "-C", _CSRC, "-j8"] + ([] if verbose else ["-s"])
    subprocess.check_call(cmd)
    build_ext(verbose)
    return LIB_PATH


def build_ext(verbose: bool = False) -> str:
    """g++ on torch_binding.cpp against this interpreter's torch; in-tree output (jspsr_b200/_jspsr_torch.so) so that
    it travels with the repository snapshot.  Rebuilt when the source, the headers or torch changed."""
    import sysconfig
    import torch
    from torch.utils import cpp_extension
    deps = [_EXT_SRC, LIB_PATH, os.path.join(_CSRC, "..", "..", "include", "jspsr_spn.h"),
            os.path.join(_CSRC, "..", "..", "include", "jspsr_tiles.h"),
            os.path.join(_CSRC, "..", "..", "include", "jspsr_peer.h"), torch.__file__]
    if os.path.exists(EXT_PATH) and all(os.path.getmtime(EXT_PATH) >= os.path.getmtime(d) for d in deps):
        return EXT_PATH
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DTORCH_EXTENSION_NAME=_jspsr_torch",
           "-DTORCH_API_INCLUDE_EXTENSION_H", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    cmd += ["-I" + d for d in cpp_extension.include_paths()]
    cmd += ["-I" + sysconfig.get_paths()["include"], "-I/usr/local/cuda/include", _EXT_SRC, "-o", EXT_PATH,
            "-L" + tlib, "-ltorch", "-ltorch_cpu", "-ltorch_cuda", "-lc10", "-lc10_cuda", "-ltorch_python",
            "-L" + _CSRC, "-ljspsr_spn", "-Wl,-rpath,$ORIGIN/csrc", "-Wl,-rpath," + tlib]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return EXT_PATH


def ext():
    """The C++ torch extension (None when it has not been built, or JSPSR_NO_TORCH_EXT=1): the same kernels of the
    same library, with the per-call bookkeeping in C++ instead of Python + ctypes."""
    global _ext, _ext_tried
    if not _ext_tried:
        _ext_tried = True
        if os.environ.get("JSPSR_NO_TORCH_EXT", "0") != "1" and os.path.exists(EXT_PATH):
            lib()                                  # libjspsr_spn.so first: the extension links against it
            import torch  # noqa: F401  (libtorch must be loaded before the extension)
            from . import _jspsr_torch
            if _jspsr_torch.abi_version() != lib().jspsr_version():
                raise RuntimeError("jspsr_b200/_jspsr_torch.so was built against another libjspsr_spn.so: rebuild")
            _ext = _jspsr_torch
    return _ext


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C jspsr_b200/csrc`). jspsr_b200 has no CPU or PyTorch fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def exported_symbols():
    return list(_SIGNATURES)


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().jspsr_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (status {rc}): {msg}")
