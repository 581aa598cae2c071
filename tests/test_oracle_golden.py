"""The oracle (oracle/gpuar_oracle.c) against the reference's golden vectors.

The vectors in tests/golden/ were produced by the reference's own arCompress
(gpuar_kernel.cu:487) -- see tests/golden/make_golden.py.  CPU only.
"""
import numpy as np
import pytest

import _oracle as O
from _vectors import SMALL, SWEEP, VECTORS, make_input, md5, stored_payload
from gpuar_b200 import datagen as D


@pytest.mark.parametrize("name", SMALL)
def test_oracle_encode_matches_reference_golden(name):
    rec = VECTORS[name]
    data = make_input(rec)
    assert data.size == rec["n"] and md5(data) == rec["input_md5"]
    pay = O.encode(data)
    assert pay.size == rec["payload_bytes"]
    assert md5(pay) == rec["payload_md5"]
    stored = stored_payload(name)
    if stored is not None:
        assert np.array_equal(pay, stored)
    assert O.index(pay).size == rec["packets"]


@pytest.mark.parametrize("name", SMALL)
def test_oracle_decode_round_trip(name):
    data = make_input(VECTORS[name])
    assert np.array_equal(O.decode(O.encode(data)), data)


@pytest.mark.parametrize("packet", sorted(int(k) for k in SWEEP))
def test_oracle_packet_size_sweep(packet):
    rec = SWEEP[str(packet)]
    data = D.mixed(rec["seed"], rec["n"])
    pay = O.encode(data, packet)
    assert pay.size == rec["payload_bytes"] and md5(pay) == rec["payload_md5"]
    assert np.array_equal(O.decode(pay), data)


def test_known_answers_spelled_out():
    # 1-byte input 00 and 'abracadabra'
    assert O.encode(np.zeros(1, np.uint8)).tobytes().hex() == "060001000040"
    assert O.encode(np.frombuffer(b"abracadabra", np.uint8)).tobytes().hex() == "0f000b006163100524151e55a17380"


def test_empty_input_is_header_only():
    assert O.encode(np.zeros(0, np.uint8)).size == 0
    g = O.gip_file(np.zeros(0, np.uint8))
    assert g.size == 20 and g[:3].tolist() == [0, 1, 0] and g[12] == 20


def test_header_layout():
    # file_header.hpp:19-36,61-72: sizes are u32 LE at bytes 4 and 12; total includes the header
    data = D.uniform(3, 20000)
    g = O.gip_file(data)
    assert int.from_bytes(g[4:8].tobytes(), "little") == 20000
    assert int.from_bytes(g[12:16].tobytes(), "little") == g.size
    h = O.header((1 << 34) + 5, (1 << 33) + 7)   # 32-bit truncation, file_header.hpp:61-72
    assert int.from_bytes(h[4:8].tobytes(), "little") == 5
    assert int.from_bytes(h[12:16].tobytes(), "little") == 7


def test_masked_compare_ignores_only_undefined_bytes():
    g = O.gip_file(D.uniform(4, 9000))
    h = g.copy()
    for k in O.HEADER_MASKED:
        h[k] ^= 0xA5
    assert O.masked_equal(g, h)
    h[4] ^= 1
    assert not O.masked_equal(g, h)


def test_sizes_and_bounds():
    # worst observed packet (round-robin / adversarial) stays below the 8704-byte slot, gpu.h:12
    for data in (D.round_robin(8192), D.adversarial_packet()):
        assert O.encode(data).size == 8281
    assert O.encode(D.zeros(8192)).size == 210


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("n", [1, 15, 16, 17, 255, 4097, 8192])
def test_oracle_equals_reference_build_ragged(n):
    for seed in (1, 2):
        for gen in (D.uniform, D.and3):
            data = gen(seed, n)
            a, b = O.encode(data), O.ref_encode(data)
            assert np.array_equal(a, b)
            assert np.array_equal(O.ref_decode(a, n), data)
