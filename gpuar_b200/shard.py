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
    """Per-rank handle.  world == 1: nothing to exchange."""

    def __init__(self, dev, rank: int, world: int, layout: str = "segments"):
        """layout: "segments" = the concatenated stream in `world` equal segments, segment g on GPU g
        (every GPU receives the same number of bytes); "gather" = the whole stream on rank 0."""
        self.dev, self.rank, self.world = dev, rank, world
        self._peer = None
        if world > 1:
            from ._peer import PeerConcat
            self._peer = PeerConcat(rank, world, layout)

    def reserve(self, cap_per_rank: int) -> None:
        if self._peer:
            self._peer.reserve(cap_per_rank)

    def concat(self, payload, total) -> None:
        """Land this rank's payload[:total] at its scanned offset in the concatenated stream."""
        if self._peer:
            self._peer.concat(payload, total)
