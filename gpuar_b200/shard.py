"""Packet-range sharding of one input across the GPUs of one box.

Packets are independent (fresh model and coder state per packet, reference
src/gpuar_kernel.cu:901-907), so rank r of W encodes packets [r*P/W, (r+1)*P/W) with no
data-path collective.  The only exchange is the W payload totals: their exclusive scan is
where each rank's stream lands in the gathered payload, which then equals the single-GPU
payload byte for byte because concatenation order = packet order.
"""
from __future__ import annotations

PACKET = 8192


def packet_range(n_bytes: int, rank: int, world: int) -> tuple[int, int]:
    """Packets [p0, p1) owned by `rank`: contiguous, balanced to within one packet."""
    packets = (n_bytes + PACKET - 1) // PACKET
    return packets * rank // world, packets * (rank + 1) // world


def byte_range(n_bytes: int, rank: int, world: int) -> tuple[int, int]:
    p0, p1 = packet_range(n_bytes, rank, world)
    return min(n_bytes, p0 * PACKET), min(n_bytes, p1 * PACKET)


def exclusive_scan(totals) -> list[int]:
    out, run = [], 0
    for t in totals:
        out.append(run)
        run += int(t)
    return out


class ShardedCodec:
    """Per-rank handle on one stream shared by ``world`` GPUs.  world == 1: the plain device codec."""

    def __init__(self, dev, rank: int, world: int, layout: str = "segments"):
        """layout: "segments" = the concatenated stream in `world` equal segments, segment g on GPU g
        (every GPU receives the same number of bytes); "gather" = the whole stream on rank 0."""
        self.dev, self.rank, self.world, self.layout = dev, rank, world, layout
        self.group = None
        self._scratch = None
        self._dscratch = None
        if world > 1:
            from ._peer import PeerGroup
            self.group = PeerGroup(rank, world, layout)

    def reserve(self, cap_per_rank: int) -> None:
        if self.group:
            self.group.reserve(cap_per_rank)

    def _buf(self, attr: str, nbytes: int):
        import torch
        cur = getattr(self, attr)
        if cur is None or cur.numel() < nbytes:
            cur = torch.empty(nbytes + 256, dtype=torch.uint8, device="cuda")
            setattr(self, attr, cur)
        return cur

    def encode(self, x) -> None:
        """This rank's packet range -> its place in the concatenated stream (gpuar_b200_encode_sharded).
        Collective over the ranks, but nothing here talks to another process: the totals travel
        through the peer-mapped mailboxes.  Layout in group.layout_out (device)."""
        import ctypes as C
        import torch
        from ._lib import check, lib
        g = self.group
        n = x.numel()
        scratch = self._buf("_scratch", int(lib().gpuar_b200_encode_scratch_bytes(n)))
        check(lib().gpuar_b200_encode_sharded(C.byref(g.shard), x.data_ptr() if n else None, n, g.layout_out.data_ptr(),
                                              None, scratch.data_ptr(), scratch.numel(),
                                              torch.cuda.current_stream().cuda_stream), "gpuar_b200_encode_sharded")

    def decode(self, stream_bytes: int, out, result) -> None:
        """Packets that start in this rank's segment -> out[k * 8192] (gpuar_b200_decode_sharded);
        result = int64[8] on the device (packets, raw bytes, status, packets before, raw bytes before)."""
        import ctypes as C
        import torch
        from ._lib import check, lib
        g = self.group
        seg = int(lib().gpuar_b200_shard_segment_bytes(stream_bytes, self.world))
        scratch = self._buf("_dscratch", int(lib().gpuar_b200_decode_sharded_scratch_bytes(seg, out.numel() // PACKET)))
        check(lib().gpuar_b200_decode_sharded(C.byref(g.shard), stream_bytes, out.data_ptr(), out.numel(),
                                              result.data_ptr(), scratch.data_ptr(), scratch.numel(),
                                              torch.cuda.current_stream().cuda_stream), "gpuar_b200_decode_sharded")
