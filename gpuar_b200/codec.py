"""Host-side plumbing over the C ABI: device memory, streams and scratch via PyTorch.

This module computes nothing itself.  Every byte of codec work happens in
libgpuar_b200.so (hand-written sm_100a kernels); PyTorch only owns the buffers.
It mirrors the reference's seam (src/gpuar.h:59-86 + what GPUCompressor does
around it, src/gpu_compressor.cpp:84-395):

    DeviceCodec.encode   device bytes   -> compacted .gip payload on the device
    DeviceCodec.index    payload        -> packet offsets (device chain discovery)
    DeviceCodec.decode   payload+offsets-> device bytes
    compress/decompress  host bytes <-> .gip image through the library's own
                         staged H2D / kernels / D2H pipeline
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import FILE_HEADER, PACKET, PAD, SLOT, GpuarError, check, lib

_inited = set()


def init(device: int | None = None) -> None:
    """gpuar_b200_init on `device` (default: current).  Raises without a usable B200."""
    if not torch.cuda.is_available():
        raise GpuarError(_lib.E_NODEVICE, "gpuar_b200 needs a CUDA device (no CPU fallback)")
    dev = torch.cuda.current_device() if device is None else int(device)
    if dev in _inited:
        return
    with torch.cuda.device(dev):
        check(lib().gpuar_b200_init(), "gpuar_b200_init")
    _inited.add(dev)


def packets_of(n: int) -> int:
    return (n + PACKET - 1) // PACKET


def payload_bound(n: int, packet: int = PACKET) -> int:
    return int(lib().gpuar_b200_payload_bound_ex(n, packet))


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class DeviceCodec:
    """Device-resident codec with cached scratch (grown on demand, reused across calls)."""

    def __init__(self, device: int | None = None):
        self.device = torch.cuda.current_device() if device is None else int(device)
        init(self.device)
        self._scratch = None
        self._iscratch = None
        self._pscratch = None

    def _buf(self, attr: str, nbytes: int) -> torch.Tensor:
        cur = getattr(self, attr)
        if cur is None or cur.numel() < nbytes:
            cur = torch.empty(nbytes + 256, dtype=torch.uint8, device=f"cuda:{self.device}")
            setattr(self, attr, cur)
        return cur

    # ------------------------------------------------------------------ encode
    def encode(self, x: torch.Tensor, payload: torch.Tensor | None = None, total: torch.Tensor | None = None,
               sizes: torch.Tensor | None = None, packet: int = PACKET):
        """x: uint8 CUDA tensor.  Returns (payload buffer, total[1] int64 on device, sizes or None).

        The payload occupies payload[:total]; nothing is synchronised.  `packet` = raw bytes per
        packet (8192 = the reference format; other multiples of 16 up to 16112 for sweeps)."""
        assert x.is_cuda and x.dtype == torch.uint8 and x.is_contiguous()
        n = x.numel()
        cap = payload_bound(n, packet)
        if payload is None:
            payload = torch.empty(cap + 16, dtype=torch.uint8, device=x.device)
        if total is None:
            total = torch.zeros(1, dtype=torch.int64, device=x.device)
        scratch = self._buf("_scratch", int(lib().gpuar_b200_encode_scratch_bytes_ex(n, packet)))
        with torch.cuda.device(self.device):
            check(lib().gpuar_b200_encode_ex(x.data_ptr(), n, packet, payload.data_ptr(), payload.numel(),
                                             total.data_ptr(), sizes.data_ptr() if sizes is not None else None,
                                             scratch.data_ptr(), scratch.numel(), _stream()), "gpuar_b200_encode_ex")
        return payload, total, sizes

    # ------------------------------------------------------------------- index
    def index(self, payload: torch.Tensor, c: int, max_packets: int, offsets: torch.Tensor | None = None,
              result: torch.Tensor | None = None, packet: int = PACKET):
        """Packet offsets of payload[:c] (payload readable PAD bytes past c).

        Returns (offsets int64[max_packets], result int64[4] = packets, raw bytes, status, ragged flag)."""
        assert payload.is_cuda and payload.dtype == torch.uint8 and payload.numel() >= c + PAD
        if offsets is None:
            offsets = torch.empty(max(1, max_packets), dtype=torch.int64, device=payload.device)
        if result is None:
            result = torch.zeros(4, dtype=torch.int64, device=payload.device)
        scratch = self._buf("_iscratch", int(lib().gpuar_b200_index_scratch_bytes(c)))
        with torch.cuda.device(self.device):
            check(lib().gpuar_b200_index_ex(payload.data_ptr(), c, packet, offsets.data_ptr(), max_packets,
                                            result.data_ptr(), scratch.data_ptr(), scratch.numel(), _stream()),
                  "gpuar_b200_index_ex")
        return offsets, result

    # ------------------------------------------------------------------ decode
    def decode(self, payload: torch.Tensor, c: int, offsets: torch.Tensor, n_packets: int,
               out: torch.Tensor | None = None, packet: int = PACKET) -> torch.Tensor:
        """Decode n_packets packets; packet p lands at out[p*packet:]."""
        if out is None:
            out = torch.empty(max(1, n_packets) * packet, dtype=torch.uint8, device=payload.device)
        with torch.cuda.device(self.device):
            check(lib().gpuar_b200_decode_ex(payload.data_ptr(), c, packet, offsets.data_ptr(), n_packets,
                                             out.data_ptr(), out.numel(), _stream()), "gpuar_b200_decode_ex")
        return out

    def decode_packed(self, payload: torch.Tensor, c: int, offsets: torch.Tensor, n_packets: int,
                      out: torch.Tensor, packet: int = PACKET) -> torch.Tensor:
        """Decode n_packets packets back to back into `out` (short packets allowed anywhere, as in
        the reference's CPU decoder).  Returns total[1] int64 on device = raw bytes of the packets;
        packets that would end past out.numel() are not written."""
        total = torch.zeros(1, dtype=torch.int64, device=payload.device)
        scratch = self._buf("_pscratch", int(lib().gpuar_b200_decode_packed_scratch_bytes(n_packets, packet)))
        with torch.cuda.device(self.device):
            check(lib().gpuar_b200_decode_packed(payload.data_ptr(), c, packet, offsets.data_ptr(), n_packets,
                                                 out.data_ptr(), out.numel(), total.data_ptr(), scratch.data_ptr(),
                                                 scratch.numel(), _stream()), "gpuar_b200_decode_packed")
        return total

    # -------------------------------------------------- convenience (synchronises)
    def encode_bytes(self, x: torch.Tensor, packet: int = PACKET) -> torch.Tensor:
        payload, total, _ = self.encode(x, packet=packet)
        return payload[: int(total.item())]

    def decode_bytes(self, payload: torch.Tensor, packet: int = PACKET) -> torch.Tensor:
        """payload: exact-length uint8 CUDA tensor (any capacity); discovers the chain, decodes."""
        c = payload.numel()
        padded = torch.zeros(c + PAD + 16, dtype=torch.uint8, device=payload.device)
        padded[:c] = payload
        max_packets = c // 5 + 1
        offsets, result = self.index(padded, c, max_packets, packet=packet)
        packets, raw, status, ragged = (int(v) for v in result.tolist())
        if status != 0:
            raise GpuarError(status, "gpuar_b200_index")
        if ragged:                                                  # short packets before the last one
            out = torch.empty(raw + 16, dtype=torch.uint8, device=payload.device)
            total = self.decode_packed(padded, c, offsets, packets, out, packet=packet)
            assert int(total.item()) == raw
            return out[:raw]
        out = self.decode(padded, c, offsets, packets, packet=packet)
        return out[:raw]


# ------------------------------------------------------------ host-buffer path
def _as_u8(a) -> np.ndarray:
    if isinstance(a, np.ndarray):
        return np.ascontiguousarray(a, dtype=np.uint8)
    return np.frombuffer(bytes(a), dtype=np.uint8)


def _device_list(devices):
    if devices is None:
        return None, 1
    arr = (C.c_int * len(devices))(*[int(d) for d in devices])
    return arr, len(devices)


def compress(data, out: np.ndarray | None = None, devices=None) -> np.ndarray:
    """.gip image (20-byte header + payload) of `data`, via gpuar_b200_compress_host[_multi].
    devices: list of CUDA device ordinals to spread the chunks over (None = the current device)."""
    init()
    src = _as_u8(data)
    cap = FILE_HEADER + payload_bound(src.size)
    if out is None:
        out = np.empty(cap, dtype=np.uint8)
    assert out.size >= cap
    n_out = C.c_size_t(0)
    devs, nd = _device_list(devices)
    check(lib().gpuar_b200_compress_host_multi(devs, nd, src.ctypes.data if src.size else None, src.size,
                                               out.ctypes.data, out.size, C.byref(n_out)), "gpuar_b200_compress_host_multi")
    return out[: n_out.value]


def link_probe(src: np.ndarray, out: np.ndarray, out_bytes: int, devices=None) -> None:
    """The host<->device transfers of :func:`compress` without the kernels (gpuar_b200_host_link_probe)."""
    init()
    devs, nd = _device_list(devices)
    check(lib().gpuar_b200_host_link_probe(devs, nd, src.ctypes.data, src.size, out.ctypes.data, min(out_bytes, out.size)),
          "gpuar_b200_host_link_probe")


def raw_size(gip) -> int:
    g = _as_u8(gip)
    raw = C.c_uint64(0)
    check(lib().gpuar_b200_gip_raw_size(g.ctypes.data, g.size, C.byref(raw)), "gpuar_b200_gip_raw_size")
    return int(raw.value)


def walk(gip) -> tuple[int, int]:
    """(packets, raw bytes) of a .gip image by hopping over its packet headers on the host."""
    g = _as_u8(gip)
    packets, raw = C.c_uint64(0), C.c_uint64(0)
    check(lib().gpuar_b200_gip_walk(g.ctypes.data, g.size, C.byref(packets), C.byref(raw)), "gpuar_b200_gip_walk")
    return int(packets.value), int(raw.value)


def decompress(gip, out: np.ndarray | None = None, out_cap: int | None = None, devices=None) -> np.ndarray:
    """Inverse of :func:`compress`, via gpuar_b200_decompress_host[_multi]."""
    init()
    g = _as_u8(gip)
    if out is None:
        if out_cap is None:
            # The header field is 32 bit in reference-written files.  Images written by this library
            # carry (and mark) the high half; for an unmarked image that COULD hold 4 GiB or more (8192
            # equal bytes code into 210, so from ~110 MB of payload on) the size comes from a hop over
            # its packet headers (one read per packet).
            out_cap = raw_size(g)
            if g.size >= FILE_HEADER and g[3] != 0xB2 and (g.size // 210 + 1) * PACKET >= (1 << 32):
                out_cap = walk(g)[1]
        out = np.empty(out_cap + PACKET, dtype=np.uint8)
    n_out = C.c_size_t(0)
    devs, nd = _device_list(devices)
    check(lib().gpuar_b200_decompress_host_multi(devs, nd, g.ctypes.data, g.size, out.ctypes.data, out.size,
                                                 C.byref(n_out)), "gpuar_b200_decompress_host_multi")
    return out[: n_out.value]


def write_header(raw_bytes: int, gip_bytes: int) -> np.ndarray:
    h = np.zeros(FILE_HEADER, dtype=np.uint8)
    lib().gpuar_b200_write_header(h.ctypes.data, raw_bytes, gip_bytes)
    return h
