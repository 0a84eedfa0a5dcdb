"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports
every symbol include/jspsr_spn.h declares; argument validation happens before any CUDA work;
the Python mirror keeps the reference's constructor / state_dict contract and refuses CPU tensors
(no fallback).  No kernel is launched here."""
import ctypes
import os
import re
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADERS = [os.path.join(ROOT, "include", h) for h in ("jspsr_spn.h", "jspsr_tiles.h", "jspsr_peer.h")]


@pytest.fixture(scope="module")
def lib():
    from jspsr_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib


def declared_functions():
    text = "".join(open(h).read() for h in HEADERS)
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(jspsr_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    names = declared_functions()
    assert len(names) == 29, names
    handle = ctypes.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} is declared in include/*.h but not exported"
    assert sorted(lib.exported_symbols()) == names, "jspsr_b200/_lib.py binds a different set than the header"


def test_peer_argument_validation_needs_no_gpu(lib):
    """include/jspsr_peer.h: the structs are validated on the host before any CUDA work."""
    from jspsr_b200 import peer
    h = lib.lib()
    one = ctypes.c_void_p(16)
    pr = peer.PeerReduceStruct()
    pr.rank, pr.world, pr.average = 0, 9, 1
    bwd = lambda pr_, gw=one: h.jspsr_spn_backward_reduce(one, one, one, one, one, None, one, one, gw, one, one, 1, 8, 8, 1, 1.0,
                                                          0, 0, ctypes.addressof(pr_), None)
    assert bwd(pr) == -1 and b"at most 8" in h.jspsr_last_error()
    pr.world, pr.rank = 2, 2
    assert bwd(pr) == -1 and b"rank 2 / world 2" in h.jspsr_last_error()
    pr.rank = 1
    assert bwd(pr) == -1 and b"slots[0]" in h.jspsr_last_error()
    pr.slots[0], pr.slots[1] = 256, 512
    assert bwd(pr, None) == -1 and b"nothing to reduce" in h.jspsr_last_error()

    sp = peer.StripPeerStruct()
    fwd = lambda: h.jspsr_spn_forward_strip_peer(one, one, one, one, one, one, 64, 128, 256, 0, 0, 72, 0, 1.0, 0, None,
                                                 ctypes.addressof(sp), None)
    assert fwd() == -1 and b"my_flags" in h.jspsr_last_error()
    sp.my_flags, sp.halo = 256, 65
    assert fwd() == -1 and b"halo 65" in h.jspsr_last_error()
    sp.halo, sp.up_dst = 8, 1024
    assert fwd() == -1 and b"without its neighbour" in h.jspsr_last_error()
    assert h.jspsr_spn_forward_strip_peer(one, one, one, one, one, one, 64, 128, 256, 0, 0, 72, 0, 1.0, 2, None,
                                          ctypes.addressof(sp), None) == -2          # mixed dtype
    assert h.jspsr_spn_forward_strip_peer(one, one, one, one, one, one, 64, 128, 256, 0, 0, 72, 0, 1.0, 0, None,
                                          None, None) == -1
    sp.up_dst, sp.up_flags = None, 2048
    assert h.jspsr_strip_halo_push(one, 64, 128, 0, ctypes.addressof(sp), None) == -1
    assert b"destination for every neighbour" in h.jspsr_last_error()
    sp.up_flags = None
    assert h.jspsr_strip_halo_push(one, 64, 128, 0, ctypes.addressof(sp), None) == 0   # a single strip: nothing to do
    assert h.jspsr_peer_open(None, None) == -1 and h.jspsr_peer_close(None) == 0 and h.jspsr_peer_free(None) == 0


def test_version_and_workspace(lib):
    h = lib.lib()
    assert h.jspsr_version() == 107
    assert 64 <= h.jspsr_spn_workspace_bytes() <= 4096
    assert h.jspsr_spn_host_scratch_bytes(2, 128, 128, 0) >= 2 * 2 * 128 * 128 * 4 * 29


def test_argument_validation_needs_no_gpu(lib):
    h = lib.lib()
    one = ctypes.c_void_p(16)  # never dereferenced: validation fails first
    cases = [
        (lambda: h.jspsr_spn_forward(one, one, one, one, one, one, 0, 8, 8, 1, 1.0, 0, None), -1, "non-positive"),
        (lambda: h.jspsr_spn_forward(one, one, one, one, one, one, 1, 8, 8, 7, 1.0, 0, None), -1, "norm_mode"),
        (lambda: h.jspsr_spn_forward(one, one, one, one, one, one, 1, 8, 8, 1, 1.0, 5, None), -1, "dtype"),
        (lambda: h.jspsr_spn_forward(None, one, one, one, one, one, 1, 8, 8, 1, 1.0, 0, None), -1, "null"),
        (lambda: h.jspsr_spn_forward(one, one, one, one, one, one, 1, 1 << 25, 8, 1, 1.0, 0, None), -2, "2^22"),
        (lambda: h.jspsr_spn_forward(ctypes.c_void_p(18), one, one, one, one, one, 1, 8, 8, 1, 1.0, 0, None), -4, "aligned"),
        (lambda: h.jspsr_gen_spn_forward(one, one, one, one, one, one, one, None, None, 1, 32, 8, 8, 1, 1.0, 0, None), -2, "C = 64"),
        (lambda: h.jspsr_gen_spn_forward(one, one, one, one, one, one, one, None, None, 1, 64, 8, 8, 1, 1.0, 1, None), -2, "fp32"),
        (lambda: h.jspsr_gen_spn_forward(one, one, one, one, one, one, one, one, None, 1, 64, 8, 8, 1, 1.0, 0, None), -1, "together"),
        (lambda: h.jspsr_gen_spn_forward(one, one, ctypes.c_void_p(20), one, one, one, one, None, None, 1, 64, 8, 8, 1, 1.0, 0, None), -4, "conv_w"),
        (lambda: h.jspsr_gen_tail_grad_params(one, one, one, one, one, 1, 32, 8, 8, 0, None), -2, "C = 64 and C = 128"),
        (lambda: h.jspsr_gen_tail_grad_params(one, one, one, one, one, 1, 64, 8, 8, 2, None), -1, "dtype is 0"),
        (lambda: h.jspsr_gen_tail_grad_params(one, one, None, None, one, 1, 64, 8, 8, 0, None), -1, "null"),
        (lambda: h.jspsr_gen_tail_grad_params(one, one, one, one, None, 1, 64, 8, 8, 0, None), -1, "null"),
        (lambda: h.jspsr_gen_tail_grad_params(one, one, one, one, ctypes.c_void_p(24), 1, 64, 8, 8, 0, None), -4, "workspace"),
        (lambda: h.jspsr_spn_iterate(one, one, one, None, None, one, None, 1, 8, 8, 2, 2, None), -2, "mixed"),
        (lambda: h.jspsr_spn_forward(one, one, one, one, one, one, 1, 8, 8, 1, 1.0, 3, None), -1, "dtype"),
        (lambda: h.jspsr_spn_iterate(one, one, one, None, None, one, None, 1, 8, 8, 0, 0, None), -1, "T="),
        (lambda: h.jspsr_spn_iterate(one, one, one, one, None, one, None, 1, 8, 8, 2, 0, None), -1, "together"),
        (lambda: h.jspsr_spn_forward_strip(one, one, one, one, one, one, 1, 8, 8, 4, 0, 0, 4, 1, 1.0, 0, None, None), -1,
         "strip"),
        (lambda: h.jspsr_nlspn_affinity_forward(one, None, one, one, one, 1, 8, 8, 9, 0, 0, None), -1, "affinity"),
        (lambda: h.jspsr_spn_backward(one, one, one, one, one, None, one, one, one, one, None, 1, 8, 8, 1, 1.0, 0, 0, None),
         -1, "workspace"),
    ]
    for call, want, needle in cases:
        rc = call()
        assert rc == want, (rc, want, needle)
        assert needle in h.jspsr_last_error().decode(), (needle, h.jspsr_last_error())


def test_preserve_blend_argument_validation_needs_no_gpu(lib):
    """jspsr_preserve_blend (the LRRU cascade's blend, LRRU.py:447-451): host-side checks come before any CUDA work."""
    h = lib.lib()
    one = ctypes.c_void_p(16)
    assert h.jspsr_preserve_blend(one, one, one, -1, 0, None) == -1 and b"negative" in h.jspsr_last_error()
    assert h.jspsr_preserve_blend(one, one, one, 8, 2, None) == -2 and b"dtype" in h.jspsr_last_error()
    assert h.jspsr_preserve_blend(None, None, None, 0, 0, None) == 0          # an empty batch is not an error
    assert h.jspsr_preserve_blend(one, None, one, 8, 0, None) == -1 and b"null" in h.jspsr_last_error()
    assert h.jspsr_preserve_blend(one, ctypes.c_void_p(18), one, 8, 0, None) == -4
    assert h.jspsr_preserve_blend(one, ctypes.c_void_p(17), one, 8, 1, None) == -4
    import jspsr_b200 as jb
    x = torch.zeros(1, 1, 4, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        jb.functional.preserve_blend(x, x)


def test_iterate_backward_argument_validation_needs_no_gpu(lib):
    """jspsr_spn_iterate_backward (autograd of the loop nlspn.py:222-235): what it does not cover is refused on the host,
    so that the caller runs T applications of jspsr_spn_backward instead."""
    h = lib.lib()
    one = ctypes.c_void_p(16)
    call = lambda T=6, dtype=0, gl=one, scratch=one, B=2: h.jspsr_spn_iterate_backward(gl, one, one, one, one, one, one, one,
                                                                                        scratch, B, 8, 8, T, dtype, None)
    assert call(T=9) == -2 and b"[1, 8]" in h.jspsr_last_error()
    assert call(T=0) == -2
    assert call(dtype=1) == -2 and b"fp32 only" in h.jspsr_last_error()
    assert call(B=0) == -1
    assert call(gl=None) == -1 and b"null" in h.jspsr_last_error()
    assert call(scratch=None) == -1                      # T > 1 needs the carry scratch
    assert call(gl=ctypes.c_void_p(18)) == -4


def test_modules_keep_the_reference_contract():
    import jspsr_b200 as jb
    pp = jb.PostProcessor(kernel_size=3, residual=True, scale=1.0)
    assert [k for k, _ in pp.named_parameters()] == ["w", "b"]
    assert tuple(pp.w.shape) == (1, 1, 3, 3) and tuple(pp.b.shape) == (1,)
    assert pp.w.requires_grad and pp.b.requires_grad and torch.all(pp.w == 1) and torch.all(pp.b == 0)
    assert (pp.stride, pp.padding, pp.dilation, pp.scale, pp.residual) == ((1, 1), (1, 1), (1, 1), 1.0, True)
    lr = jb.Post_process_deconv(types.SimpleNamespace(kernel_size=3, dkn_residual=True))
    assert set(lr.state_dict()) == {"w", "b"} and lr.dkn_residual is True and lr.im2col_step == 64
    args = types.SimpleNamespace(prop_time=6, affinity="TGASS", affinity_gamma=0.5, conf_prop=True,
                                 preserve_input=False, legacy=False)
    nl = jb.NLSPN(args, 8, 1, 3, 3)
    assert set(nl.state_dict()) == {"conv_offset_aff.weight", "conv_offset_aff.bias", "aff_scale_const", "w", "b", "w_conf"}
    assert float(nl.aff_scale_const) == 4.0 and nl.aff_scale_const.requires_grad
    assert not nl.w.requires_grad and not nl.b.requires_grad and not nl.w_conf.requires_grad
    assert torch.all(nl.conv_offset_aff.weight == 0) and torch.all(nl.conv_offset_aff.bias == 0)
    assert tuple(nl.conv_offset_aff.weight.shape) == (24, 8, 3, 3)
    tc = jb.NLSPN(types.SimpleNamespace(prop_time=1, affinity="TC", affinity_gamma=0.5, conf_prop=False,
                                        preserve_input=False, legacy=False), 8, 1, 3, 3)
    assert float(tc.aff_scale_const) == 8.0 and not tc.aff_scale_const.requires_grad
    with pytest.raises(NotImplementedError):
        jb.PostProcessor(kernel_size=5)
    with pytest.raises(AssertionError):
        jb.NLSPN(args, 8, 2, 3, 3)


def test_reference_checkpoint_keys_load():
    """utils/utils.py:360-364 keeps a checkpoint entry only if key AND shape match the model's."""
    import jspsr_b200 as jb

    class Host(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.postprocessor = jb.PostProcessor(3, True, 1.0)

    ckpt = {"postprocessor.w": torch.full((1, 1, 3, 3), 0.5), "postprocessor.b": torch.full((1,), 0.25)}
    m = Host()
    own = m.state_dict()
    kept = {k: v for k, v in ckpt.items() if k in own and v.shape == own[k].shape}
    assert set(kept) == set(ckpt)
    m.load_state_dict(kept)
    assert float(m.postprocessor.b) == 0.25
    # optimizer groups select the layer by the substring "postprocessor" (utils/common_config.py:250-253)
    assert [n for n, _ in m.named_parameters() if "postprocessor" in n] == ["postprocessor.w", "postprocessor.b"]


def test_no_cpu_fallback():
    import jspsr_b200 as jb
    from jspsr_b200 import functional as F
    pp = jb.PostProcessor()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pp(torch.rand(1, 1, 8, 8), torch.rand(1, 9, 8, 8), torch.rand(1, 18, 8, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        F.spn_iterate(torch.rand(1, 1, 8, 8), torch.rand(1, 9, 8, 8), torch.rand(1, 18, 8, 8), 2)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "jspsr_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("the oracle", ""), f"{fn} mentions the oracle package"
            assert "torchvision" not in [ln.split()[1] if ln.startswith(("import ", "from ")) else "" for ln in src.splitlines()]


def test_torch_extension_builds_and_binds_the_same_library(lib):
    """csrc/torch_binding.cpp: C++ autograd wrappers over the C ABI (no kernels of its own).  Without a GPU it must
    import, report the library's ABI version and refuse CPU tensors."""
    import torch
    lib.build_ext()
    e = lib.ext()
    assert e is not None and e.abi_version() == lib.lib().jspsr_version()
    for name in ("propagate", "spn_forward", "multi_loss", "launch_count"):
        assert hasattr(e, name)
    with pytest.raises(RuntimeError):
        e.spn_forward(torch.zeros(1, 1, 4, 4), torch.zeros(1, 9, 4, 4), torch.zeros(1, 18, 4, 4), torch.ones(1, 1, 3, 3),
                      torch.zeros(1), 1, 1.0)
    assert e.launch_count() == 0
