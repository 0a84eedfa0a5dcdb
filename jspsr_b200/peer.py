"""Peer-mapped device memory between the ranks of one NVSwitch node (include/jspsr_peer.h).

One process per GPU.  Each rank allocates a buffer through the library (cudaMalloc + CUDA IPC handle),
`torch.distributed` carries the 64-byte handles to the other ranks (plumbing only), every rank maps the
others' buffers, and from then on the KERNELS talk to each other directly over NVLink: the gradient
all-reduce inside `spn_backward_kernel` and the halo exchange inside `spn_forward_kernel` (strips.py).
New relative to the reference, which is single-process.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional

import torch
import torch.distributed as dist

from . import _lib

MAX_RANKS = 8          # JSPSR_PEER_MAX_RANKS
HANDLE_BYTES = 64      # JSPSR_PEER_HANDLE_BYTES
REDUCE_BYTES = 2 * MAX_RANKS * 16 * 8 + 64   # JSPSR_REDUCE_BYTES
STRIP_FLAG_BYTES = 64  # JSPSR_STRIP_FLAG_BYTES


class PeerReduceStruct(ctypes.Structure):
    """jspsr_peer_reduce"""
    _fields_ = [("slots", ctypes.c_void_p * MAX_RANKS), ("rank", ctypes.c_int), ("world", ctypes.c_int),
                ("average", ctypes.c_int)]


class StripPeerStruct(ctypes.Structure):
    """jspsr_strip_peer"""
    _fields_ = [("up_dst", ctypes.c_void_p), ("dn_dst", ctypes.c_void_p), ("up_flags", ctypes.c_void_p),
                ("dn_flags", ctypes.c_void_p), ("my_flags", ctypes.c_void_p), ("stamp", ctypes.c_uint),
                ("halo", ctypes.c_int)]


class _DevicePointer:
    """Lets torch view library-owned device memory as a tensor (CUDA array interface, no copy)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PeerMemory:
    """`nbytes` of zero-filled device memory on every rank of `group`, each rank's buffer mapped into all the others.

    ptrs[r] is rank r's buffer as addressable from THIS process (ptrs[rank] is the local allocation).
    `only` restricts the mapping to some ranks (row strips only need their two neighbours)."""

    def __init__(self, nbytes: int, group=None, only: Optional[List[int]] = None, device: Optional[torch.device] = None):
        if not dist.is_initialized():
            raise RuntimeError("PeerMemory needs an initialised torch.distributed process group (one process per GPU)")
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.nbytes = int(nbytes)
        lib = _lib.lib()
        local = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(HANDLE_BYTES)
        with torch.cuda.device(self.device):
            _lib.check(lib.jspsr_peer_alloc(self.nbytes, ctypes.byref(local), handle), "jspsr_peer_alloc")
        self.local_ptr = int(local.value)
        handles: List[Optional[bytes]] = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        self.ptrs: List[Optional[int]] = [None] * self.world
        self._opened: List[int] = []
        with torch.cuda.device(self.device):
            for r in range(self.world):
                if r == self.rank:
                    self.ptrs[r] = self.local_ptr
                elif only is None or r in only:
                    p = ctypes.c_void_p()
                    _lib.check(lib.jspsr_peer_open(handles[r], ctypes.byref(p)), f"jspsr_peer_open(rank {r})")
                    self.ptrs[r] = int(p.value)
                    self._opened.append(int(p.value))
        dist.barrier(group=group)   # nobody uses (or frees) a buffer before everyone has mapped it
        self._closed = False

    def tensor(self, dtype=torch.uint8, offset_bytes: int = 0, numel: Optional[int] = None) -> torch.Tensor:
        """The LOCAL buffer as a torch tensor (a view: it must not outlive this object)."""
        t = torch.as_tensor(_DevicePointer(self.local_ptr, self.nbytes), device=self.device)
        t = t[offset_bytes:]
        if numel is not None:
            t = t[:numel * torch.empty((), dtype=dtype).element_size()]
        return t.view(dtype)

    def close(self) -> None:
        """Collective: unmap the peers' buffers, then free the local one."""
        if self._closed:
            return
        self._closed = True
        lib = _lib.lib()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            for p in self._opened:
                lib.jspsr_peer_close(ctypes.c_void_p(p))
            if dist.is_initialized():
                dist.barrier(group=self.group)
            lib.jspsr_peer_free(ctypes.c_void_p(self.local_ptr))


class PeerGradReducer:
    """All-reduce of PostProcessor.w / .b gradients fused into the backward kernel (jspsr_spn_backward_reduce).

    Attach with `postprocessor.set_grad_reducer(reducer)`; every rank must then run the same sequence of backward
    passes (the step stamp is a device-side counter).  `average=True` is DistributedDataParallel's convention; with
    DDP around the model, exclude the two parameters from its buckets
    (`DistributedDataParallel._set_params_and_buffers_to_ignore_for_model(model, ["postprocessor.w", "postprocessor.b"])`)
    so they are reduced once, here."""

    def __init__(self, group=None, average: bool = True):
        self.mem = PeerMemory(REDUCE_BYTES, group)
        if self.mem.world > MAX_RANKS:
            raise RuntimeError(f"PeerGradReducer supports up to {MAX_RANKS} ranks (one NVSwitch node)")
        self.rank, self.world = self.mem.rank, self.mem.world
        self.struct = PeerReduceStruct()
        for r in range(self.world):
            self.struct.slots[r] = self.mem.ptrs[r]
        self.struct.rank, self.struct.world, self.struct.average = self.rank, self.world, int(bool(average))

    @property
    def address(self) -> int:
        """Host address of the jspsr_peer_reduce struct (what the C ABI takes)."""
        return ctypes.addressof(self.struct)

    def close(self) -> None:
        self.mem.close()
