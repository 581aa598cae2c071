"""NVLink concatenation of the per-rank payloads (one process per GPU, torch.distributed plumbing).

The concatenated stream (rank order = packet order) is laid out in `world` equal segments,
segment g on GPU g (`layout="segments"`, default), or entirely on rank 0 (`layout="gather"`).
Setup (once): every owner allocates its segment with the library's cudaMalloc wrapper and exports
a CUDA IPC handle; the handles are all-gathered and mapped by the other ranks (the mapping enables
peer access over NVLink).  Per step: the ranks' 8-byte totals are all-gathered on the device and
each rank launches gpuar_b200_shard_concat, whose kernel derives the segment size and its own
landing offset from the totals and writes its stream with peer stores.  No size ever visits the
host, and no payload byte goes through NCCL.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from ._lib import check, lib


class PeerConcat:
    def __init__(self, rank: int, world: int, layout: str = "segments"):
        assert layout in ("segments", "gather")
        self.rank, self.world, self.layout = rank, world, layout
        self.n_segments = world if layout == "segments" else 1
        self.ptrs = None            # device pointers of the segments, valid in THIS process
        self.mine = None            # the allocation this rank owns (or None)
        self.cap = 0
        self.totals = torch.zeros(world, dtype=torch.int64, device="cuda")
        self.layout_out = torch.zeros(3, dtype=torch.int64, device="cuda")   # total, segment size, my base

    def reserve(self, cap_per_rank: int) -> None:
        # a segment holds ceil(total / n_segments) <= the largest per-rank bound (+ rounding)
        cap = int(cap_per_rank) + 4096 if self.layout == "segments" else int(cap_per_rank) * self.world
        if self.ptrs is not None and cap <= self.cap:
            return
        self.release()
        owner = self.rank < self.n_segments
        handle = torch.zeros(64, dtype=torch.uint8, device="cuda")
        if owner:
            ptr = C.c_void_p()
            check(lib().gpuar_b200_device_alloc(cap, C.byref(ptr)), "gpuar_b200_device_alloc")
            buf = (C.c_uint8 * 64)()
            check(lib().gpuar_b200_ipc_export(ptr, buf), "gpuar_b200_ipc_export")
            handle.copy_(torch.frombuffer(bytearray(buf), dtype=torch.uint8))
            self.mine = ptr.value
        handles = torch.zeros(self.world * 64, dtype=torch.uint8, device="cuda")
        dist.all_gather_into_tensor(handles, handle)
        handles = handles.cpu().numpy().reshape(self.world, 64)
        ptrs = []
        for g in range(self.n_segments):
            if g == self.rank:
                ptrs.append(self.mine)
            else:
                raw = (C.c_uint8 * 64).from_buffer_copy(handles[g].tobytes())
                ptr = C.c_void_p()
                check(lib().gpuar_b200_ipc_open(raw, C.byref(ptr)), "gpuar_b200_ipc_open")
                ptrs.append(ptr.value)
        self.ptrs = ptrs
        self.c_ptrs = (C.c_void_p * self.n_segments)(*ptrs)
        self.cap = cap
        dist.barrier()

    def concat(self, payload: torch.Tensor, total: torch.Tensor) -> None:
        dist.all_gather_into_tensor(self.totals, total)
        check(lib().gpuar_b200_shard_concat(payload.data_ptr(), self.totals.data_ptr(), self.rank, self.world,
                                            self.c_ptrs, self.n_segments, self.cap, self.layout_out.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream), "gpuar_b200_shard_concat")

    def my_segment(self):
        """(tensor view of this rank's segment, valid bytes in it) after a concat; synchronises."""
        torch.cuda.synchronize()
        total, seg, _ = (int(v) for v in self.layout_out.tolist())
        if self.mine is None:
            return None, 0
        if self.layout == "gather":
            valid = total
        else:
            valid = max(0, min(seg, total - self.rank * seg))
        return torch.as_tensor(_RawDeviceBuffer(self.mine, max(valid, 1)), device="cuda")[:valid], valid

    def release(self) -> None:
        if self.ptrs is None:
            return
        torch.cuda.synchronize()
        for g, p in enumerate(self.ptrs):
            if g == self.rank:
                lib().gpuar_b200_device_free(p)
            else:
                lib().gpuar_b200_ipc_close(p)
        self.ptrs = None
        self.mine = None
        self.cap = 0


class _RawDeviceBuffer:
    """__cuda_array_interface__ view of a raw device pointer, so torch can wrap it without a copy."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}
