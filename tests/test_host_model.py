"""Closed forms used by the CUDA kernels (gpuar_b200/csrc/coder_math.h, encode_math.h, decode_math.h), compiled for the
host and driven through a lane-by-lane emulation of the kernels' data flow
(tests/host_model.cpp), against the oracle and the reference's golden vectors.  CPU only."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import _oracle as O
from _vectors import SMALL, VECTORS, make_input, md5
from gpuar_b200 import datagen as D

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libhost_model.so")


@pytest.fixture(scope="module")
def model():
    src = os.path.join(HERE, "host_model.cpp")
    hdrs = [os.path.join(HERE, "..", "gpuar_b200", "csrc", h) for h in ("coder_math.h", "decode_math.h", "encode_math.h")]
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(f) for f in [src] + hdrs):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", src, "-o", SO])
    lib = C.CDLL(SO)
    lib.host_model_encode_stream.restype = C.c_size_t
    lib.host_model_encode_stream.argtypes = [O._u8p, C.c_size_t, O._u8p, C.c_uint32]
    lib.host_model_encode_stream_ws.restype = C.c_size_t
    lib.host_model_encode_stream_ws.argtypes = [O._u8p, C.c_size_t, O._u8p, C.c_uint32]
    lib.host_model_carry_events.restype = C.c_uint64
    lib.host_model_decode_packet.restype = C.c_uint32
    lib.host_model_decode_packet.argtypes = [O._u8p, C.c_size_t, C.c_size_t, O._u8p, C.c_int]
    lib.host_model_check_division.restype = C.c_uint64
    lib.host_model_check_division.argtypes = [C.c_uint32, C.c_uint32]
    lib.host_model_check_renorm.restype = C.c_uint64
    lib.host_model_check_renorm.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32]
    lib.host_model_check_unscale.restype = C.c_uint64
    lib.host_model_check_unscale.argtypes = [C.c_uint32, C.c_uint32]
    return lib


def model_encode(lib, data, packet=8192, ws=False):
    """encode_kernel's lane (ws=False) or the stages of encode_ws_kernel (ws=True)"""
    buf = np.zeros(O.n_packets(data.size, packet) * (packet + 512) + 64, np.uint8)
    src = data if data.size else np.zeros(1, np.uint8)
    fn = lib.host_model_encode_stream_ws if ws else lib.host_model_encode_stream
    return buf[: fn(O._ptr(src), data.size, O._ptr(buf), packet)].copy()


def model_decode(lib, pay, n, latency=False):
    """decode_kernel's lane: the throughput step or the latency step (decode_math.h)"""
    c = pay.size
    padded = np.zeros((c + 64 + 15) // 16 * 16, np.uint8)
    padded[:c] = pay
    out = np.zeros(n + O.PACKET, np.uint8)
    pos = 0
    for o in O.index(pay):
        pos += lib.host_model_decode_packet(O._ptr(padded), c + 64, int(o), out[pos:].ctypes.data_as(O._u8p), int(latency))
    return out[:pos]


def both_decoders_return(lib, pay, data):
    for latency in (False, True):
        assert np.array_equal(model_decode(lib, pay, data.size, latency), data)


def test_reciprocal_division_is_exact(model):
    # floor(n / T) for every total T = 256..8447 around every multiple of T up to the largest numerator,
    # and for the largest packet size the format admits (16112: totals up to 16367, numerators < 2^30)
    assert model.host_model_check_division(8448 * 65536, 8192) == 0
    assert model.host_model_check_division(16368 * 65536, 16112) == 0


def test_closed_form_renormalisation_equals_reference_loop(model):
    # 4M random reachable coder states per packet size against the bit-at-a-time loop
    assert model.host_model_check_renorm(1, 4_000_000, 8192) == 0
    assert model.host_model_check_renorm(2, 2_000_000, 16112) == 0


def test_float_estimated_divide_is_exact(model):
    assert model.host_model_check_unscale(13, 8192) == 0
    assert model.host_model_check_unscale(29, 16112) == 0


@pytest.mark.parametrize("name", SMALL)
def test_kernel_math_matches_reference_golden(model, name):
    rec = VECTORS[name]
    data = make_input(rec)
    pay = model_encode(model, data)
    assert pay.size == rec["payload_bytes"] and md5(pay) == rec["payload_md5"]
    assert np.array_equal(model_encode(model, data, ws=True), pay)
    both_decoders_return(model, pay, data)


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 63, 64, 65, 4095, 8191, 8192])
def test_kernel_math_ragged_lengths(model, n):
    for data in (D.uniform(n, n), D.and3(n + 1, n), D.zeros(n), D.round_robin(n)):
        pay = model_encode(model, data)
        assert np.array_equal(pay, O.encode(data))
        assert np.array_equal(model_encode(model, data, ws=True), pay)
        both_decoders_return(model, pay, data)


@pytest.mark.parametrize("packet", [4096, 12288, 16112])
def test_kernel_math_other_packet_sizes(model, packet):
    from _vectors import SWEEP
    rec = SWEEP[str(packet)]
    data = D.mixed(rec["seed"], rec["n"])
    pay = model_encode(model, data, packet)
    assert pay.size == rec["payload_bytes"] and md5(pay) == rec["payload_md5"]
    assert np.array_equal(model_encode(model, data, packet, ws=True), pay)
    both_decoders_return(model, pay, data)


def test_kernel_math_long_underflow_runs(model):
    # two-symbol inputs straddling the midpoint keep the coder in the 01../10.. state; the greedy straddle
    # packet drives the reference's pending-underflow counter into the thousands
    rng = np.random.default_rng(5)
    carries_before = model.host_model_carry_events()
    straddle, max_pend = D.straddle_packet(8192, report=True)
    assert max_pend > 1000
    for k in range(9):
        data = straddle if k == 8 else rng.choice(np.array([127, 128], np.uint8), size=8192, p=[0.5, 0.5])
        pay = model_encode(model, data)
        assert np.array_equal(pay, O.encode(data))
        assert np.array_equal(model_encode(model, data, ws=True), pay)
        both_decoders_return(model, pay, data)
    # the encoder had to carry into words it had already stored (the kernels' rare path) on these inputs
    assert model.host_model_carry_events() > carries_before


def test_kernel_math_random_distributions(model):
    # seeded sweep over byte distributions the fixed vectors do not hold: heavy skew (a few symbols take almost all
    # the mass, so counts run high and intervals get narrow), runs, ramps, two-symbol alphabets at both ends of the
    # byte range, random lengths around the packet size
    rng = np.random.default_rng(20261017)
    for k in range(48):
        n = int(rng.integers(1, 3 * 8192 + 1))
        kind = k % 6
        if kind == 0:
            p = rng.dirichlet(np.full(256, 0.02))
            data = rng.choice(256, size=n, p=p).astype(np.uint8)
        elif kind == 1:
            data = np.repeat(rng.integers(0, 256, size=n // 37 + 1, dtype=np.uint8), 37)[:n]
        elif kind == 2:
            data = (np.arange(n) // int(rng.integers(1, 65))).astype(np.uint8)
        elif kind == 3:
            data = rng.choice(np.array([0, 255], np.uint8), size=n, p=[0.97, 0.03])
        elif kind == 4:
            data = np.minimum(rng.geometric(0.3, size=n) - 1, 255).astype(np.uint8)
        else:
            data = np.sort(rng.integers(0, 256, size=n, dtype=np.uint8))
        pay = model_encode(model, data)
        assert np.array_equal(pay, O.encode(data)), (k, n)
        assert np.array_equal(model_encode(model, data, ws=True), pay), (k, n)
        both_decoders_return(model, pay, data)
