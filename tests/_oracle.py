"""ctypes bindings to the checkers: oracle/libgpuar_oracle.so (C restatement) and,
when present, oracle/_ref/libgpuar_ref.so (the reference's own codec compiled
from /root/reference by `make -C oracle ref`).

Test infrastructure only -- the product (gpuar_b200/) never imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
PORT_SO = os.path.join(ORACLE_DIR, "libgpuar_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libgpuar_ref.so")

PACKET = 8192
SLOT = 8704
HEADER = 20
HEADER_MASKED = (3, 8, 9, 10, 11, 16, 17, 18, 19)   # bytes the reference never writes

_u8p = C.POINTER(C.c_uint8)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_u8p)


def build_port():
    src = os.path.join(ORACLE_DIR, "gpuar_oracle.c")
    if not os.path.exists(PORT_SO) or os.path.getmtime(PORT_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "port"], stdout=subprocess.DEVNULL)
    return PORT_SO


_port = None


def port():
    global _port
    if _port is None:
        lib = C.CDLL(build_port())
        lib.gpuar_oracle_encode_packet.restype = C.c_size_t
        lib.gpuar_oracle_encode_packet.argtypes = [_u8p, C.c_size_t, _u8p]
        lib.gpuar_oracle_decode_packet.restype = C.c_size_t
        lib.gpuar_oracle_decode_packet.argtypes = [_u8p, _u8p]
        lib.gpuar_oracle_encode_stream.restype = C.c_size_t
        lib.gpuar_oracle_encode_stream.argtypes = [_u8p, C.c_size_t, _u8p, C.c_size_t]
        lib.gpuar_oracle_decode_stream.restype = C.c_size_t
        lib.gpuar_oracle_decode_stream.argtypes = [_u8p, C.c_size_t, _u8p, C.c_size_t]
        lib.gpuar_oracle_index.restype = C.c_size_t
        lib.gpuar_oracle_index.argtypes = [_u8p, C.c_size_t, C.POINTER(C.c_uint64), C.c_size_t]
        lib.gpuar_oracle_write_header.restype = None
        lib.gpuar_oracle_write_header.argtypes = [_u8p, C.c_uint64, C.c_uint64]
        _port = lib
    return _port


def have_ref() -> bool:
    return os.path.exists(REF_SO)


_ref = None


def ref():
    global _ref
    if _ref is None:
        lib = C.CDLL(REF_SO)
        lib.gpuar_ref_encode_packet.restype = C.c_size_t
        lib.gpuar_ref_encode_packet.argtypes = [_u8p, C.c_size_t, _u8p]
        lib.gpuar_ref_decode_packet.restype = C.c_size_t
        lib.gpuar_ref_decode_packet.argtypes = [_u8p, _u8p]
        lib.gpuar_ref_encode_stream.restype = C.c_size_t
        lib.gpuar_ref_encode_stream.argtypes = [_u8p, C.c_size_t, _u8p, C.c_int]
        lib.gpuar_ref_decode_stream.restype = C.c_size_t
        lib.gpuar_ref_decode_stream.argtypes = [_u8p, C.c_size_t, _u8p, C.c_int]
        lib.gpuar_ref_gpu_init.restype = None
        lib.gpuar_ref_gpu_encode.restype = None
        lib.gpuar_ref_gpu_encode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        lib.gpuar_ref_gpu_decode.restype = None
        lib.gpuar_ref_gpu_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        lib.gpuar_ref_gpu_sync.restype = C.c_int
        _ref = lib
    return _ref


def n_packets(n: int, packet: int = PACKET) -> int:
    return (n + packet - 1) // packet


# ------------------------------------------------------------------ port API
def encode(data, packet: int = PACKET) -> np.ndarray:
    """Payload (.gip bytes from offset 20) of ``data`` by the C restatement."""
    data = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data)
    buf = np.zeros(n_packets(data.size, packet) * (packet + 512) + 16, dtype=np.uint8)
    src = data if data.size else np.zeros(1, dtype=np.uint8)
    c = port().gpuar_oracle_encode_stream(_ptr(src), data.size, _ptr(buf), packet)
    return buf[:c].copy()


def decode(payload, out_cap: int | None = None) -> np.ndarray:
    payload = np.ascontiguousarray(payload, dtype=np.uint8)
    padded = np.zeros(payload.size + 16, dtype=np.uint8)
    padded[: payload.size] = payload
    if out_cap is None:
        out_cap = raw_size(payload)
    out = np.zeros(out_cap + 16, dtype=np.uint8)
    n = port().gpuar_oracle_decode_stream(_ptr(padded), payload.size, _ptr(out), out_cap)
    if n == C.c_size_t(-1).value:
        raise ValueError("malformed packet chain")
    return out[:n].copy()


def index(payload) -> np.ndarray:
    payload = np.ascontiguousarray(payload, dtype=np.uint8)
    padded = np.zeros(payload.size + 16, dtype=np.uint8)
    padded[: payload.size] = payload
    cap = payload.size // 5 + 1
    offs = np.zeros(cap, dtype=np.uint64)
    k = port().gpuar_oracle_index(_ptr(padded), payload.size, offs.ctypes.data_as(C.POINTER(C.c_uint64)), cap)
    if k == C.c_size_t(-1).value:
        raise ValueError("malformed packet chain")
    return offs[:k].copy()


def raw_size(payload) -> int:
    """Sum of rawLen over the chain."""
    offs = index(payload).astype(np.int64)
    p = np.asarray(payload, dtype=np.uint8).astype(np.int64)
    return int((p[offs + 2] | (p[offs + 3] << 8)).sum()) if offs.size else 0


def header(raw_bytes: int, gip_bytes: int) -> np.ndarray:
    h = np.zeros(HEADER, dtype=np.uint8)
    port().gpuar_oracle_write_header(_ptr(h), raw_bytes, gip_bytes)
    return h


def gip_file(data, packet: int = PACKET) -> np.ndarray:
    """Full .gip image (header + payload) as the reference DEFINES it; undefined bytes zero."""
    pay = encode(data, packet)
    n = len(data)
    return np.concatenate([header(n, HEADER + pay.size), pay])


def masked_equal(a: np.ndarray, b: np.ndarray) -> bool:
    """.gip equality under the reference's header mask."""
    if a.size != b.size or a.size < HEADER:
        return False
    a = a.copy()
    b = b.copy()
    for k in HEADER_MASKED:
        a[k] = 0
        b[k] = 0
    return bool(np.array_equal(a, b))


# ------------------------------------------------------------- reference API
def ref_encode(data, threads: int = 1) -> np.ndarray:
    data = np.ascontiguousarray(data, dtype=np.uint8)
    buf = np.zeros(n_packets(data.size) * SLOT + 16, dtype=np.uint8)
    src = data if data.size else np.zeros(1, dtype=np.uint8)
    c = ref().gpuar_ref_encode_stream(_ptr(src), data.size, _ptr(buf), threads)
    return buf[:c].copy()


def ref_decode(payload, out_cap: int, threads: int = 1) -> np.ndarray:
    payload = np.ascontiguousarray(payload, dtype=np.uint8)
    out = np.zeros(out_cap + PACKET, dtype=np.uint8)
    n = ref().gpuar_ref_decode_stream(_ptr(payload), payload.size, _ptr(out), threads)
    return out[:n].copy()
