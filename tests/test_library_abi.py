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
                           (_lib.OPT_WS_MAX_PACKETS, (0, 23680), (-1,))):
        for v in bad:
            assert _lib.lib().gpuar_b200_set_option(key, v) == _lib.E_ARG, (key, v)
        for v in good:
            assert _lib.lib().gpuar_b200_set_option(key, v) == 0, (key, v)
    assert _lib.lib().gpuar_b200_set_option(99, 0) == _lib.E_ARG
