"""The C-ABI library loads and exports every symbol include/gpuar_b200.h declares (CPU only:
no compute calls), and its host-only helpers behave like the reference's FileHeader."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import _oracle as O
from gpuar_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "gpuar_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"^[A-Za-z_][\w \*]*?\b(\w+)\s*\([^;{]*\)\s*;", text, flags=re.M)
    return sorted(set(names))


def test_header_declares_what_the_binding_knows():
    assert sorted(_lib.SIGNATURES) == declared_functions()


def test_library_exports_every_declared_symbol():
    handle = _lib.lib()
    for name in declared_functions():
        assert getattr(handle, name) is not None, name


def test_constants_match_reference_format():
    # gpu.h:12-14, file_header.hpp:19-22
    assert (_lib.PACKET, _lib.SLOT, _lib.FILE_HEADER) == (8192, 8704, 20)
    h = _lib.lib()
    assert h.gpuar_b200_abi_version() == 1
    assert h.gpuar_b200_packets(0) == 0 and h.gpuar_b200_packets(1) == 1 and h.gpuar_b200_packets(8193) == 2
    assert h.gpuar_b200_payload_bound(8192 * 3) == 3 * 8704 + _lib.PAD
    assert h.gpuar_b200_encode_scratch_bytes(1 << 20) >= 128 * 8704
    assert h.gpuar_b200_index_scratch_bytes(1 << 20) > 0


def test_header_writer_matches_oracle_in_defined_bytes():
    h = np.zeros(20, np.uint8)
    for raw, gip in ((0, 20), (12345, 6789), ((1 << 32) + 7, (1 << 33) + 9)):
        _lib.lib().gpuar_b200_write_header(h.ctypes.data, raw, gip)
        ref = O.header(raw, gip)
        for k in range(20):
            if k not in O.HEADER_MASKED:
                assert h[k] == ref[k]
        assert _lib.lib().gpuar_b200_check_header(h.ctypes.data) == 0
        # 64-bit extension lives only in bytes the reference leaves undefined
        assert int.from_bytes(h[4:12].tobytes(), "little") == raw
        assert int.from_bytes(h[12:20].tobytes(), "little") == gip
    h[1] = 2
    assert _lib.lib().gpuar_b200_check_header(h.ctypes.data) == _lib.E_FORMAT


def test_raw_size_ignores_garbage_in_undefined_bytes():
    g = O.gip_file(np.arange(20000, dtype=np.uint32).astype(np.uint8))
    raw = C.c_uint64()
    assert _lib.lib().gpuar_b200_gip_raw_size(g.ctypes.data, g.size, C.byref(raw)) == 0 and raw.value == 20000
    g[8:12] = 0xEE                      # what a reference-written file may contain there
    assert _lib.lib().gpuar_b200_gip_raw_size(g.ctypes.data, g.size, C.byref(raw)) == 0 and raw.value == 20000
    assert _lib.lib().gpuar_b200_gip_raw_size(g.ctypes.data, 10, C.byref(raw)) == _lib.E_FORMAT


def test_no_device_means_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    assert _lib.lib().gpuar_b200_init() == _lib.E_NODEVICE
    n = C.c_size_t()
    src = np.zeros(8192, np.uint8)
    dst = np.zeros(20 + 8704 + 64, np.uint8)
    rc = _lib.lib().gpuar_b200_compress_host(src.ctypes.data, src.size, dst.ctypes.data, dst.size, C.byref(n))
    assert rc != 0
    from gpuar_b200 import codec
    with pytest.raises(_lib.GpuarError):
        codec.compress(src)


def test_strerror():
    assert _lib.strerror(0) == "ok"
    assert "argument" in _lib.strerror(_lib.E_ARG)


def test_options_validate_their_values():
    """gpuar_b200_set_option: unknown keys and out-of-range values are refused (no device needed)."""
    for key, good, bad in ((_lib.OPT_ENCODE_PATH, (0, 1, 2), (-1, 3)),
                           (_lib.OPT_COMPACT_TILE, (4, 8, 16, 32, 64, 128, 0), (-4, 2, 3, 24, 256)),
                           (_lib.OPT_WS_MAX_PACKETS, (0, 23680), (-1,)),
                           (_lib.OPT_DECODE_PATH, (1, 2, 0), (-1, 3, 1 << 32))):
        for v in bad:
            assert _lib.lib().gpuar_b200_set_option(key, v) == _lib.E_ARG, (key, v)
        for v in good:
            assert _lib.lib().gpuar_b200_set_option(key, v) == 0, (key, v)
    assert _lib.lib().gpuar_b200_set_option(99, 0) == _lib.E_ARG


def test_wide_sizes_are_trusted_only_under_the_mark():
    """Header byte 3 = GPUAR_HEADER_WIDE_MARK says bytes 8-11 / 16-19 carry the high halves of 64-bit sizes;
    without it (reference-written files leave those bytes uninitialised, file_header.hpp:28-36,61-72) only the
    32-bit fields count -- a stray 1 in byte 8 must not turn 20000 into 4 GiB + 20000."""
    data = np.arange(20000, dtype=np.uint32).astype(np.uint8)
    g = O.gip_file(data)
    raw = C.c_uint64()
    h = np.zeros(20, np.uint8)
    _lib.lib().gpuar_b200_write_header(h.ctypes.data, data.size, g.size)
    assert h[3] == 0xB2
    g[:20] = h
    g[8] = 1                                               # marked, but implausible for this payload: ignored
    assert _lib.lib().gpuar_b200_gip_raw_size(g.ctypes.data, g.size, C.byref(raw)) == 0 and raw.value == 20000
    g[3] = 0                                               # unmarked: ignored whatever it says
    assert _lib.lib().gpuar_b200_gip_raw_size(g.ctypes.data, g.size, C.byref(raw)) == 0 and raw.value == 20000
    # a header alone (the CLI passes 20 bytes and the file size): plausible wide size under the mark is taken
    big_raw, big_gip = (5 << 30) + 123, (3 << 30) + 20
    _lib.lib().gpuar_b200_write_header(h.ctypes.data, big_raw, big_gip)
    assert _lib.lib().gpuar_b200_gip_raw_size(h.ctypes.data, big_gip, C.byref(raw)) == 0 and raw.value == big_raw
    h[3] = 0x5A
    assert _lib.lib().gpuar_b200_gip_raw_size(h.ctypes.data, big_gip, C.byref(raw)) == 0 and raw.value == big_raw % (1 << 32)


def test_gip_walk_matches_the_oracle_chain():
    """gpuar_b200_gip_walk: the host hop over compLen / rawLen (cpu_compressor.cpp:47-78), no device needed."""
    from gpuar_b200 import datagen as D
    packets, raw = C.c_uint64(), C.c_uint64()
    for data in (np.zeros(0, np.uint8), D.mixed(3, 8192 * 5 + 77), D.uniform(9, 1)):
        g = O.gip_file(data)
        assert _lib.lib().gpuar_b200_gip_walk(g.ctypes.data, g.size, C.byref(packets), C.byref(raw)) == 0
        assert (packets.value, raw.value) == ((data.size + 8191) // 8192, data.size)
    # short packets in the middle: legal for the CPU decoder
    parts = [D.uniform(1, 8192 + 5), D.and3(2, 100)]
    pay = np.concatenate([O.encode(p) for p in parts])
    g = np.concatenate([O.header(0, 20 + pay.size), pay])
    assert _lib.lib().gpuar_b200_gip_walk(g.ctypes.data, g.size, C.byref(packets), C.byref(raw)) == 0
    assert (packets.value, raw.value) == (3, 8192 + 5 + 100)
    # broken chains
    bad = g.copy()
    bad[20] = 3; bad[21] = 0                               # compLen 3 <= header length
    assert _lib.lib().gpuar_b200_gip_walk(bad.ctypes.data, bad.size, C.byref(packets), C.byref(raw)) == _lib.E_FORMAT
    assert _lib.lib().gpuar_b200_gip_walk(g.ctypes.data, g.size - 1, C.byref(packets), C.byref(raw)) == _lib.E_FORMAT
    bad = g.copy()
    bad[22] = 0x01; bad[23] = 0x20                         # rawLen 8193
    assert _lib.lib().gpuar_b200_gip_walk(bad.ctypes.data, bad.size, C.byref(packets), C.byref(raw)) == _lib.E_UNSUPPORTED


def test_segment_layout_of_a_sharded_stream():
    """gpuar_b200_shard_segment_bytes: equal segments, 256-byte granules, never smaller than a packet with its halo."""
    f = _lib.lib().gpuar_b200_shard_segment_bytes
    assert f(0, 8) == 16384 and f(5055, 8) == 16384
    for total, n in ((10_380_379_762, 8), (67_648_274 * 8, 8), (1 << 20, 3), (1 << 40, 16)):
        s = f(total, n)
        assert s % 256 == 0 and s * n >= total and (s - 256) * n < total or s == 16384
    assert f(1 << 30, 1) == 1 << 30 and f(100, 0) == 0
    assert _lib.lib().gpuar_b200_decode_sharded_scratch_bytes(1 << 20, 200) >= 200 * 8 + _lib.lib().gpuar_b200_index_scratch_bytes(1 << 20)


def test_sharded_entry_points_refuse_malformed_groups():
    """argument checks of the sharded ABI come before any device work"""
    sh = _lib.Shard()
    lay = (C.c_uint64 * 8)()
    for rank, world, nseg in ((0, 0, 1), (2, 2, 2), (0, 17, 17), (0, 2, 3)):
        sh.rank, sh.world, sh.n_segments = rank, world, nseg
        assert _lib.lib().gpuar_b200_encode_sharded(C.byref(sh), None, 0, lay, None, lay, 0, None) == _lib.E_ARG
        assert _lib.lib().gpuar_b200_decode_sharded(C.byref(sh), 0, None, 0, lay, lay, 0, None) == _lib.E_ARG
    devs = (C.c_int * 2)(0, 0)
    n = C.c_size_t()
    buf = np.zeros(20 + 8704 + 64, np.uint8)
    assert _lib.lib().gpuar_b200_compress_host_multi(devs, 0, buf.ctypes.data, 1, buf.ctypes.data, buf.size, C.byref(n)) != 0
