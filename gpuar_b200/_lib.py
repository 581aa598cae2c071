"""ctypes binding of libgpuar_b200.so (the C ABI in include/gpuar_b200.h).

The library is the product: if it is missing or cannot run on this device, every
call raises -- there is no Python, PyTorch or CPU fallback for the codec.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# GPUAR_B200_LIB points at an alternative build of the same library (A/B measurements)
LIB_PATH = os.environ.get("GPUAR_B200_LIB") or os.path.join(HERE, "libgpuar_b200.so")

PACKET = 8192          # GPUAR_PACKET_BYTES  (reference gpu.h:13)
SLOT = 8704            # GPUAR_SLOT_BYTES    (reference gpu.h:12)
FILE_HEADER = 20       # GPUAR_FILE_HEADER   (reference file_header.hpp:19-22)
PAD = 64               # GPUAR_PAD_BYTES

E_ARG, E_FORMAT, E_NODEVICE, E_UNSUPPORTED = -1, -2, -3, -4


class GpuarError(RuntimeError):
    def __init__(self, code: int, what: str):
        self.code = code
        super().__init__(f"{what}: {strerror(code)} ({code})")


_u8p = C.POINTER(C.c_uint8)
_vp = C.c_void_p
_sz = C.c_size_t

# name -> (restype, argtypes); kept in the order of include/gpuar_b200.h
SIGNATURES = {
    "gpuar_b200_abi_version": (C.c_int, []),
    "gpuar_b200_strerror": (C.c_char_p, [C.c_int]),
    "gpuar_b200_init": (C.c_int, []),
    "gpuar_b200_packets": (_sz, [_sz]),
    "gpuar_b200_payload_bound": (_sz, [_sz]),
    "gpuar_b200_encode_scratch_bytes": (_sz, [_sz]),
    "gpuar_b200_index_scratch_bytes": (_sz, [_sz]),
    "gpuar_b200_encode": (C.c_int, [_vp, _sz, _vp, _sz, _vp, _vp, _vp, _sz, _vp]),
    "gpuar_b200_index": (C.c_int, [_vp, _sz, _vp, _sz, _vp, _vp, _sz, _vp]),
    "gpuar_b200_decode": (C.c_int, [_vp, _sz, _vp, _sz, _vp, _sz, _vp]),
    "gpuar_b200_decode_packed_scratch_bytes": (_sz, [_sz, _sz]),
    "gpuar_b200_decode_packed": (C.c_int, [_vp, _sz, _sz, _vp, _sz, _vp, _sz, _vp, _vp, _sz, _vp]),
    "gpuar_b200_payload_bound_ex": (_sz, [_sz, _sz]),
    "gpuar_b200_encode_scratch_bytes_ex": (_sz, [_sz, _sz]),
    "gpuar_b200_encode_ex": (C.c_int, [_vp, _sz, _sz, _vp, _sz, _vp, _vp, _vp, _sz, _vp]),
    "gpuar_b200_index_ex": (C.c_int, [_vp, _sz, _sz, _vp, _sz, _vp, _vp, _sz, _vp]),
    "gpuar_b200_decode_ex": (C.c_int, [_vp, _sz, _sz, _vp, _sz, _vp, _sz, _vp]),
    "gpuar_b200_compress_host": (C.c_int, [_vp, _sz, _vp, _sz, C.POINTER(_sz)]),
    "gpuar_b200_decompress_host": (C.c_int, [_vp, _sz, _vp, _sz, C.POINTER(_sz)]),
    "gpuar_b200_compress_host_multi": (C.c_int, [C.POINTER(C.c_int), C.c_int, _vp, _sz, _vp, _sz, C.POINTER(_sz)]),
    "gpuar_b200_decompress_host_multi": (C.c_int, [C.POINTER(C.c_int), C.c_int, _vp, _sz, _vp, _sz, C.POINTER(_sz)]),
    "gpuar_b200_host_link_probe": (C.c_int, [C.POINTER(C.c_int), C.c_int, _vp, _sz, _vp, _sz]),
    "gpuar_b200_gip_raw_size": (C.c_int, [_vp, _sz, C.POINTER(C.c_uint64)]),
    "gpuar_b200_gip_walk": (C.c_int, [_vp, _sz, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "gpuar_b200_write_header": (None, [_vp, C.c_uint64, C.c_uint64]),
    "gpuar_b200_check_header": (C.c_int, [_vp]),
    "gpuar_b200_peer_concat": (C.c_int, [_vp, C.c_int, _sz, _vp, C.c_int, _sz, _vp]),
    "gpuar_b200_encode_sharded": (C.c_int, [_vp, _vp, _sz, _vp, _vp, _vp, _sz, _vp]),
    "gpuar_b200_shard_segment_bytes": (C.c_uint64, [C.c_uint64, C.c_int]),
    "gpuar_b200_decode_sharded_scratch_bytes": (_sz, [C.c_uint64, _sz]),
    "gpuar_b200_decode_sharded": (C.c_int, [_vp, C.c_uint64, _vp, _sz, _vp, _vp, _sz, _vp]),
    "gpuar_b200_enable_peer": (C.c_int, [C.c_int]),
    "gpuar_b200_device_alloc": (C.c_int, [_sz, C.POINTER(_vp)]),
    "gpuar_b200_device_free": (C.c_int, [_vp]),
    "gpuar_b200_host_alloc": (C.c_int, [_sz, C.POINTER(_vp)]),
    "gpuar_b200_host_free": (C.c_int, [_vp]),
    "gpuar_b200_device_count": (C.c_int, []),
    "gpuar_b200_set_device": (C.c_int, [C.c_int]),
    "gpuar_b200_ipc_export": (C.c_int, [_vp, _vp]),
    "gpuar_b200_ipc_open": (C.c_int, [_vp, C.POINTER(_vp)]),
    "gpuar_b200_ipc_close": (C.c_int, [_vp]),
    "initConstantRange": (None, []),
    "garCompressExecutor": (None, [_vp, _sz, _vp, C.c_uint32]),
    "garDecompressExecutor": (None, [_vp, _sz, _vp, C.c_uint32]),
    "gpuar_b200_set_option": (C.c_int, [C.c_int, C.c_longlong]),
    "gpuar_b200_profile": (None, [C.c_int]),
    "gpuar_b200_profile_read": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "gpuar_b200_selfcheck": (C.c_int, [C.POINTER(C.c_uint64)]),
    "gpuar_b200_launch_count": (C.c_uint64, []),
}

MAX_RANKS = 16            # GPUAR_MAX_RANKS
MAILBOX_BYTES = 1024      # GPUAR_MAILBOX_BYTES


class Shard(C.Structure):
    """gpuar_b200_shard (include/gpuar_b200.h)."""
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("n_segments", C.c_int32), ("reserved", C.c_int32),
                ("seg_cap", C.c_uint64), ("segments", _vp * MAX_RANKS), ("mailbox", _vp * MAX_RANKS),
                ("calls", C.c_uint64 * 4)]


_lib = None


def lib() -> C.CDLL:
    """Load the library (once).  Raises if it has not been built -- see __graft_entry__.build()."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `make -C gpuar_b200/csrc` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`); there is no fallback path")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def strerror(code: int) -> str:
    return lib().gpuar_b200_strerror(int(code)).decode()


def check(code: int, what: str) -> None:
    if code != 0:
        raise GpuarError(code, what)


OPT_ENCODE_PATH, OPT_WS_MAX_PACKETS, OPT_COMPACT_TILE, OPT_DECODE_PATH = 1, 2, 3, 4
ENCODE_AUTO, ENCODE_FUSED, ENCODE_WS = 0, 1, 2


def set_option(key: int, value: int) -> None:
    check(lib().gpuar_b200_set_option(key, value), "gpuar_b200_set_option")


SPANS = ("encode", "compact", "index", "decode")


def profile(enable: bool) -> None:
    lib().gpuar_b200_profile(1 if enable else 0)


def profile_read() -> dict:
    """{"encode": (ms, spans), ...} accumulated since the last read (synchronises the spans)."""
    ms = (C.c_double * 4)()
    calls = (C.c_uint64 * 4)()
    check(lib().gpuar_b200_profile_read(ms, calls), "gpuar_b200_profile_read")
    return {name: (float(ms[k]), int(calls[k])) for k, name in enumerate(SPANS)}


def selfcheck() -> int:
    """Mismatches of the device self-check of the decoder's quotient (0 = it holds everywhere)."""
    bad = C.c_uint64(1)
    check(lib().gpuar_b200_selfcheck(C.byref(bad)), "gpuar_b200_selfcheck")
    return int(bad.value)


def launch_count() -> int:
    return int(lib().gpuar_b200_launch_count())
