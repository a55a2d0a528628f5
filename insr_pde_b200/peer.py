"""Peer memory for the data-parallel exchange: one IPC-exported device allocation per rank that every
rank of the box maps, and the library's own one-shot reduction kernels on top of it (``csrc/peer_kernels.cuh``) -- the
flat gradient of an iteration (3.6-56 KB for the script configurations) is read straight out of the other GPUs' memory
through NVLink / NVSwitch; no communication library sits on the data path.  torch.distributed is used ONCE, at set-up, to
hand the 64-byte IPC handles round (``all_gather_object``) and to agree that every rank could map every buffer.

``PeerBuffer.create`` returns None when peer memory cannot be set up on this box (a rank cannot map a peer: GPUs hidden
from each other by per-rank CUDA_VISIBLE_DEVICES, no peer access) -- on EVERY rank, so that all ranks fall back to the NCCL
all-reduce together.  ``INSR_PEER_ALLREDUCE=0`` switches it off.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from . import _lib, _ops

HEADER_BYTES = 4096
HEADER_FLOATS = HEADER_BYTES // 4
MAX_WORLD = 16


class _Raw:
    """a raw device range as a __cuda_array_interface__ object (zero-copy view for torch.as_tensor)"""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


def enabled():
    return os.environ.get("INSR_PEER_ALLREDUCE", "1") != "0"


class PeerBuffer:
    """``data``: this rank's ``n`` fp32 values (a torch view of the peer allocation, behind its header); ``allreduce_into``
    and ``fused.DeviceOptimizer.update_peer`` reduce it over the ranks of ``group``."""

    def __init__(self, n, device, group, base, bases, rank, world):
        self.n, self.device, self.group = int(n), device, group
        self.base, self.bases, self.rank, self.world = base, bases, rank, world
        self.bytes = HEADER_BYTES + (4 * self.n + 255) // 256 * 256
        self._raw = _Raw(base + HEADER_BYTES, max(self.n, 1))
        self.data = torch.as_tensor(self._raw, device=device)[: self.n]
        self.closed = False

    @classmethod
    def create(cls, n, device, group=None):
        """collective over ``group`` (every rank calls it with the same n); None on every rank if any rank failed"""
        if not (enabled() and dist.is_available() and dist.is_initialized()):
            return None
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world < 2 or world > MAX_WORLD:
            return None
        lib = _lib.get_lib()
        base, handle, err = None, None, None
        with _ops._DeviceGuard(device):
            try:
                base, handle = lib.peer_alloc(4 * int(n))
            except _lib.InsrError as e:
                err = str(e)
        handles = [None] * world
        dist.all_gather_object(handles, handle, group=group)
        bases = [None] * world
        if err is None and all(h is not None for h in handles):
            with _ops._DeviceGuard(device):
                try:
                    for r in range(world):
                        bases[r] = base if r == rank else lib.peer_open(handles[r])
                except _lib.InsrError as e:
                    err = str(e)
        else:
            err = err or "a peer could not allocate"
        oks = [None] * world
        dist.all_gather_object(oks, err is None, group=group)
        if not all(oks):
            with _ops._DeviceGuard(device):
                for r, b in enumerate(bases):
                    if b is not None and r != rank:
                        try:
                            lib.peer_close(b)
                        except _lib.InsrError:
                            pass
                dist.barrier(group=group)            # nobody maps this rank's allocation any more
                if base is not None:
                    lib.peer_free(base)
            if os.environ.get("INSR_PATCH_VERBOSE", "0") == "1" and err:
                import sys
                sys.stderr.write(f"[insr peer rank {rank}] peer memory unavailable ({err}); NCCL all-reduce instead\n")
            return None
        return cls(n, device, group, base, bases, rank, world)

    def allreduce_into(self, out, scale=1.0, offset=0, n=None):
        """out[:n] = scale * sum over ranks of data[offset : offset + n]  (one kernel; ``out`` is an ordinary tensor)"""
        n = self.n - offset if n is None else n
        if offset % 4 or offset < 0 or offset + n > self.n:
            raise ValueError("peer all-reduce: offset must be a multiple of 4 floats and the range inside the buffer")
        if out.device != self.data.device or out.dtype != torch.float32 or not out.is_contiguous() or out.numel() < n:
            raise ValueError("peer all-reduce: out must be a contiguous fp32 tensor on the buffer's device")
        with _ops._DeviceGuard(self.device):
            _lib.get_lib().peer_allreduce(self.world, self.rank, self.bases, HEADER_FLOATS + offset, n, float(scale),
                                          out.data_ptr(), _ops._stream(self.device))
        return out

    def healthy(self, reset=False):
        """False if a barrier inside one of the kernels gave up waiting for its peers (synchronises the device)"""
        with _ops._DeviceGuard(self.device):
            return _lib.get_lib().peer_status(self.base, reset) == 0

    def close(self):
        """collective: unmap the peers' allocations, then release the own one"""
        if self.closed:
            return
        self.closed = True
        lib = _lib.get_lib()
        with _ops._DeviceGuard(self.device):
            torch.cuda.synchronize(self.device)
            self.data = None
            if dist.is_initialized():
                dist.barrier(group=self.group)
            for r, b in enumerate(self.bases):
                if r != self.rank:
                    lib.peer_close(b)
            if dist.is_initialized():
                dist.barrier(group=self.group)
            lib.peer_free(self.base)
