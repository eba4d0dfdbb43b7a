"""Exchange buffers of the fused pmean("device") + optimiser kernel (csrc/peer.cu).

One process per GPU: every rank allocates ONE exchange buffer through the C ABI
(``mava_peer_alloc``: cudaMalloc + CUDA IPC handle), the 64-byte handles travel through
``torch.distributed.all_gather_object`` (plumbing), and every rank maps the buffers of its peers
(``mava_peer_open``).  The loss kernels write their gradients straight into the rank's own buffer
(``group.grad`` is a torch view of it); inside the optimiser kernel
(``native.reduce_clip_adam_pair``) every rank then pushes its vector over NVLink into its slot of
every peer's receive area -- the tail of the same buffer -- as 128-byte lines tagged with the call
number, and sums what it received in rank order (``MAVA_PEER_PULL=1``: flag handshake + reads of the
peers' vectors instead, the checker).  No NCCL call is left on the update path, so the CUDA graph of
an update holds no collective and the process group tears down normally.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check


class _DeviceMemory:
    """Raw device memory presented through ``__cuda_array_interface__`` (zero-copy torch view)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False),
                                         "version": 2}


class PeerGroup:
    """This rank's exchange buffer + its mappings of the peers' buffers."""

    def __init__(self, n_grad: int, device: torch.device, rank: int = 0, world: int = 1,
                 local_bufs: Optional[List[int]] = None):
        lib = _lib.load()
        self.n_grad, self.rank, self.world, self.device = int(n_grad), rank, world, device
        if world > _lib.PeerGroup._fields_[2][1]._length_:
            raise ValueError(f"at most 8 ranks per node are supported, got {world}")
        self.nbytes = int(lib.mava_peer_buffer_bytes(self.n_grad))
        self._own = C.c_void_p()
        self._opened: List[C.c_void_p] = []
        self.struct = _lib.PeerGroup()
        self.struct.rank, self.struct.world = rank, world
        if local_bufs is not None:  # all "ranks" live in this process (tests): no IPC
            self._own = None
            for r, p in enumerate(local_bufs):
                self.struct.buf[r] = p
            own_ptr = local_bufs[rank]
        else:
            handle = (C.c_ubyte * 64)()
            with torch.cuda.device(device):
                check(lib.mava_peer_alloc(self.nbytes, C.byref(self._own), handle), "mava_peer_alloc")
                own_ptr = self._own.value
                if world > 1:
                    handles: List[Optional[bytes]] = [None] * world
                    dist.all_gather_object(handles, bytes(handle))
                    for r in range(world):
                        if r == rank:
                            self.struct.buf[r] = own_ptr
                            continue
                        p = C.c_void_p()
                        hb = (C.c_ubyte * 64).from_buffer_copy(handles[r])
                        check(lib.mava_peer_open(hb, C.byref(p)), "mava_peer_open")
                        self._opened.append(p)
                        self.struct.buf[r] = p.value
                    dist.barrier()  # nobody signals before every mapping exists
                else:
                    self.struct.buf[0] = own_ptr
        raw = torch.as_tensor(_DeviceMemory(own_ptr, self.nbytes), device=device)
        self._raw = raw
        self.grad = raw[: self.n_grad * 4].view(torch.float32)  # [actor | critic | 8 loss scalars]

    @staticmethod
    def local_group(n_grad: int, device: torch.device, world: int) -> List["PeerGroup"]:
        """``world`` ranks inside ONE process on one GPU (protocol tests): plain allocations."""
        lib = _lib.load()
        nbytes = int(lib.mava_peer_buffer_bytes(n_grad))
        ptrs = []
        with torch.cuda.device(device):
            for _ in range(world):
                p = C.c_void_p()
                check(lib.mava_peer_alloc(nbytes, C.byref(p), None), "mava_peer_alloc")
                ptrs.append(p.value)
        groups = [PeerGroup(n_grad, device, r, world, local_bufs=ptrs) for r in range(world)]
        groups[0]._local_ptrs = ptrs  # freed with the first group
        return groups

    def status(self):
        """(calls completed, error word) of this rank's buffer; synchronises the current stream."""
        seq, err = C.c_uint32(), C.c_uint32()
        check(_lib.load().mava_peer_status(C.c_void_p(self.struct.buf[self.rank]), self.n_grad,
                                           C.byref(seq), C.byref(err),
                                           C.c_void_p(torch.cuda.current_stream().cuda_stream)),
              "mava_peer_status")
        return int(seq.value), int(err.value)

    def check(self) -> None:
        if self.world > 1:
            _, err = self.status()
            if err:
                raise RuntimeError("mava_reduce_clip_adam_pair: a peer handshake timed out (a rank "
                                   "died or the ranks issued different call sequences)")

    def release(self) -> None:
        """Unmap the peers' buffers and free this rank's (after the device has drained)."""
        lib = _lib.load()
        torch.cuda.synchronize(self.device)
        if self.world > 1 and self._own is not None and dist.is_initialized():
            dist.barrier()  # nobody unmaps while a peer may still read
        self.grad = self._raw = None
        for p in self._opened:
            lib.mava_peer_close(p)
        self._opened = []
        if self._own is not None and self._own.value:
            lib.mava_peer_free(self._own)
            self._own = C.c_void_p()
        for p in getattr(self, "_local_ptrs", []):
            lib.mava_peer_free(C.c_void_p(p))
        self._local_ptrs = []
