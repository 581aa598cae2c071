"""The group of GPUs that shares one concatenated stream (one process per GPU; torch.distributed
is used for the SETUP only).

Setup (once per capacity): every rank allocates its segment of the stream and its mailbox with the
library's cudaMalloc wrapper and exports CUDA IPC handles; the handles are all-gathered and mapped
by the other ranks (the mapping enables peer access over NVLink).  The result is a
``gpuar_b200_shard`` structure per rank (include/gpuar_b200.h).

Per step nothing here talks to another process: ``gpuar_b200_encode_sharded`` exchanges the ranks'
totals through the peer-mapped mailboxes and its compaction kernel writes every packet to its
final place with peer stores; ``gpuar_b200_decode_sharded`` hands the packet chain from segment to
segment the same way.  No size visits the host, no NCCL call, no payload byte is copied twice.

layout "segments": the stream lives in ``world`` equal segments, segment g on GPU g;
layout "gather":   the whole stream on rank 0.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from ._lib import MAILBOX_BYTES, Shard, check, lib

HALO = 8704 + 512          # bytes of the next segment a rank keeps behind its own (decode side)


def _export(ptr) -> torch.Tensor:
    buf = (C.c_uint8 * 64)()
    check(lib().gpuar_b200_ipc_export(ptr, buf), "gpuar_b200_ipc_export")
    return torch.frombuffer(bytearray(buf), dtype=torch.uint8)


def _open(row) -> int:
    raw = (C.c_uint8 * 64).from_buffer_copy(row.tobytes())
    ptr = C.c_void_p()
    check(lib().gpuar_b200_ipc_open(raw, C.byref(ptr)), "gpuar_b200_ipc_open")
    return ptr.value


class PeerGroup:
    def __init__(self, rank: int, world: int, layout: str = "segments"):
        assert layout in ("segments", "gather")
        self.rank, self.world, self.layout = rank, world, layout
        self.n_segments = world if layout == "segments" else 1
        self.shard = None           # gpuar_b200_shard of this rank
        self.mine = None            # the segment this rank owns (or None)
        self._mapped = []           # peer pointers to close
        self._owned = []            # own allocations to free
        self.cap = 0
        self.layout_out = torch.zeros(8, dtype=torch.int64, device="cuda")   # total, segment size, my base, my bytes, status

    def reserve(self, cap_per_rank: int) -> None:
        """Segments for a stream of at most ``cap_per_rank`` bytes from every rank."""
        # every rank must agree on the capacity (shards differ by up to a packet): take the maximum
        want = torch.tensor([int(cap_per_rank)], dtype=torch.int64, device="cuda")
        dist.all_reduce(want, op=dist.ReduceOp.MAX)
        per_rank = int(want.item())
        # a segment holds ceil(total / n_segments) <= the largest per-rank bound (+ rounding), plus
        # the head of the next segment for the decode side
        cap = (max(per_rank + 4096, 16384) + HALO if self.layout == "segments" else per_rank * self.world + HALO) + 256
        cap = (cap + 255) & ~255
        if self.shard is not None and cap <= self.cap:
            return
        self.release()
        owner = self.rank < self.n_segments
        handles = torch.zeros(2, 64, dtype=torch.uint8)
        box = C.c_void_p()
        check(lib().gpuar_b200_device_alloc(MAILBOX_BYTES, C.byref(box)), "gpuar_b200_device_alloc")
        self._owned.append(box.value)
        torch.as_tensor(_RawDeviceBuffer(box.value, MAILBOX_BYTES), device="cuda").zero_()
        handles[1] = _export(box)
        if owner:
            seg = C.c_void_p()
            check(lib().gpuar_b200_device_alloc(cap, C.byref(seg)), "gpuar_b200_device_alloc")
            self._owned.append(seg.value)
            self.mine = seg.value
            handles[0] = _export(seg)
        torch.cuda.synchronize()
        every = torch.zeros(self.world * 128, dtype=torch.uint8, device="cuda")
        dist.all_gather_into_tensor(every, handles.reshape(-1).cuda())
        every = every.cpu().numpy().reshape(self.world, 2, 64)
        sh = Shard()
        sh.rank, sh.world, sh.n_segments, sh.seg_cap = self.rank, self.world, self.n_segments, cap - HALO
        for g in range(self.world):
            if g == self.rank:
                sh.mailbox[g] = box.value
                if g < self.n_segments:
                    sh.segments[g] = self.mine
                continue
            sh.mailbox[g] = _open(every[g, 1])
            self._mapped.append(sh.mailbox[g])
            if g < self.n_segments:
                sh.segments[g] = _open(every[g, 0])
                self._mapped.append(sh.segments[g])
        self.shard = sh
        self.cap = cap
        dist.barrier()               # every mailbox is zeroed and mapped before the first call

    def segment_view(self, nbytes: int) -> torch.Tensor:
        """uint8 view of the first ``nbytes`` of this rank's own segment (no copy)."""
        assert self.mine is not None
        return torch.as_tensor(_RawDeviceBuffer(self.mine, max(nbytes, 1)), device="cuda")[:nbytes]

    def my_segment(self):
        """(tensor view of this rank's part of the stream, valid bytes in it) after an encode; synchronises."""
        torch.cuda.synchronize()
        total, seg, _, _, status = (int(v) for v in self.layout_out.tolist()[:5])
        assert status == 0, f"sharded encode status {status}"
        if self.mine is None:
            return None, 0
        valid = total if self.layout == "gather" else max(0, min(seg, total - self.rank * seg))
        return self.segment_view(valid), valid

    def release(self) -> None:
        if self.shard is None and not self._owned:
            return
        torch.cuda.synchronize()
        for p in self._mapped:
            lib().gpuar_b200_ipc_close(p)
        for p in self._owned:
            lib().gpuar_b200_device_free(p)
        self._mapped, self._owned = [], []
        self.shard = None
        self.mine = None
        self.cap = 0


class _RawDeviceBuffer:
    """__cuda_array_interface__ view of a raw device pointer, so torch can wrap it without a copy."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}
