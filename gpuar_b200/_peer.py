"""NVLink concatenation of the per-rank payloads (one process per GPU, torch.distributed plumbing).

Setup (once): rank 0 allocates the gathered payload with the library's cudaMalloc wrapper,
exports a CUDA IPC handle, the handle is broadcast, every other rank maps it (peer access
over NVLink is enabled by the mapping).  Per step: the ranks' 8-byte totals are all-gathered
on the device and each rank launches gpuar_b200_shard_concat, whose kernel computes its own
landing offset from the totals and writes its stream into rank 0's buffer with peer stores.
No size ever visits the host, and no payload byte goes through NCCL.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from ._lib import check, lib


class PeerConcat:
    def __init__(self, rank: int, world: int):
        self.rank, self.world = rank, world
        self.base = None            # gathered payload: device pointer valid in THIS process
        self.cap = 0
        self.totals = torch.zeros(world, dtype=torch.int64, device="cuda")

    def reserve(self, cap_per_rank: int) -> None:
        cap = int(cap_per_rank) * self.world
        if self.base is not None and cap <= self.cap:
            return
        self.release()
        handle = torch.zeros(64, dtype=torch.uint8, device="cuda")
        if self.rank == 0:
            ptr = C.c_void_p()
            check(lib().gpuar_b200_device_alloc(cap, C.byref(ptr)), "gpuar_b200_device_alloc")
            buf = (C.c_uint8 * 64)()
            check(lib().gpuar_b200_ipc_export(ptr, buf), "gpuar_b200_ipc_export")
            handle.copy_(torch.frombuffer(bytearray(buf), dtype=torch.uint8))
            self.base = ptr.value
        dist.broadcast(handle, src=0)
        if self.rank != 0:
            raw = (C.c_uint8 * 64).from_buffer_copy(bytes(handle.cpu().numpy().tobytes()))
            ptr = C.c_void_p()
            check(lib().gpuar_b200_ipc_open(raw, C.byref(ptr)), "gpuar_b200_ipc_open")
            self.base = ptr.value
        self.cap = cap
        dist.barrier()

    def concat(self, payload: torch.Tensor, total: torch.Tensor) -> None:
        dist.all_gather_into_tensor(self.totals, total)
        check(lib().gpuar_b200_shard_concat(payload.data_ptr(), self.totals.data_ptr(), self.rank, self.world,
                                            self.base, self.cap, torch.cuda.current_stream().cuda_stream),
              "gpuar_b200_shard_concat")

    def gathered(self, nbytes: int) -> torch.Tensor:
        """Rank 0 only: the first nbytes of the gathered payload, viewed in place as a torch tensor."""
        assert self.rank == 0
        return torch.as_tensor(_RawDeviceBuffer(self.base, nbytes), device="cuda")

    def release(self) -> None:
        if self.base is None:
            return
        torch.cuda.synchronize()
        if self.rank == 0:
            lib().gpuar_b200_device_free(self.base)
        else:
            lib().gpuar_b200_ipc_close(self.base)
        self.base = None
        self.cap = 0


class _RawDeviceBuffer:
    """__cuda_array_interface__ view of a raw device pointer, so torch can wrap it without a copy."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}
