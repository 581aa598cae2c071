"""Parity of the CUDA path (through the C ABI) with the oracle and the reference's golden
vectors.  Bit-exact: this is integer/byte work.  Run on the B200 box: pytest -m gpu."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

import _oracle as O
from _vectors import LARGE, SMALL, VECTORS, make_input, md5
from gpuar_b200 import datagen as D

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def codec():
    from gpuar_b200 import codec as cd
    cd.init()
    return cd


@pytest.fixture(scope="module")
def dev(codec):
    return codec.DeviceCodec()


def to_dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def dev_encode(dev, data):
    x = to_dev(data) if data.size else torch.empty(0, dtype=torch.uint8, device="cuda")
    return dev.encode_bytes(x).cpu().numpy()


# ------------------------------------------------------------------ encode
@pytest.mark.parametrize("name", SMALL)
def test_encode_equals_reference_golden(dev, name):
    rec = VECTORS[name]
    data = make_input(rec)
    pay = dev_encode(dev, data)
    assert pay.size == rec["payload_bytes"]
    assert md5(pay) == rec["payload_md5"]
    assert np.array_equal(pay, O.encode(data))


@pytest.mark.parametrize("n", [1, 2, 15, 16, 17, 31, 32, 33, 255, 4095, 8191, 8192, 8193, 8192 * 32,
                               8192 * 32 + 1, 8192 * 33 - 1, 8192 * 64 + 77, 8192 * 97 + 4097])
def test_encode_ragged_sizes(dev, n):
    for data in (D.uniform(n, n), D.and3(n + 3, n)):
        assert np.array_equal(dev_encode(dev, data), O.encode(data))


def test_encode_empty(dev, codec):
    assert dev_encode(dev, np.zeros(0, np.uint8)).size == 0
    g = codec.compress(np.zeros(0, np.uint8))
    assert g.size == 20 and O.masked_equal(g, O.gip_file(np.zeros(0, np.uint8)))


def test_encode_packet_sizes_output(dev):
    data = D.mixed(9, 98304)
    x = to_dev(data)
    sizes = torch.zeros(12, dtype=torch.int32, device="cuda")
    payload, total, _ = dev.encode(x, sizes=sizes)
    offs = O.index(O.encode(data))
    want = np.diff(np.append(offs, int(total.item()))).astype(np.int32)
    assert np.array_equal(sizes.cpu().numpy(), want)


def straddle_stream():
    # packets whose pending-underflow counter runs into the thousands (datagen.straddle_packet) between ordinary
    # ones, so that the lanes of a warp hit the encoders' rare paths at different steps, some not at all
    rng = np.random.default_rng(5)
    st = D.straddle_packet(8192)
    parts = []
    for p in range(70):
        if p % 3 == 0:
            parts.append(st)
        elif p % 7 == 3:
            parts.append(np.resize(st[: 8192 - 7 * p], 8192))      # the straddle run restarts inside the packet
        elif p % 3 == 1:
            parts.append(rng.choice(np.array([127, 128], np.uint8), size=8192))
        else:
            parts.append(D.uniform(p, 8192))
    return np.concatenate(parts + [st[:4097]])


@pytest.mark.parametrize("path", ["auto", "fused", "ws"])
def test_encode_long_underflow_runs(dev, path):
    from gpuar_b200 import _lib
    rng = np.random.default_rng(5)
    _lib.set_option(_lib.OPT_ENCODE_PATH, {"auto": 0, "fused": 1, "ws": 2}[path])
    try:
        data = rng.choice(np.array([127, 128], np.uint8), size=8192 * 40, p=[0.5, 0.5])
        assert np.array_equal(dev_encode(dev, data), O.encode(data))
        data = straddle_stream()
        assert np.array_equal(dev_encode(dev, data), O.encode(data))
    finally:
        _lib.set_option(_lib.OPT_ENCODE_PATH, 0)


@pytest.mark.parametrize("path", ["fused", "ws"])
def test_both_encode_kernels_are_bit_exact(dev, path):
    """The lane=packet kernel and the warp-specialised kernel must produce identical slots/payloads."""
    from gpuar_b200 import _lib
    _lib.set_option(_lib.OPT_ENCODE_PATH, _lib.ENCODE_FUSED if path == "fused" else _lib.ENCODE_WS)
    try:
        for n in (1, 31, 32, 33, 64, 65, 8191, 8192, 8193, 8192 * 31 + 5, 8192 * 32, 8192 * 70 + 4097):
            for data in (D.uniform(n + 1, n), D.and3(n + 2, n), D.zeros(n)):
                assert np.array_equal(dev_encode(dev, data), O.encode(data)), (path, n)
        rng = np.random.default_rng(7)
        data = rng.choice(np.array([127, 128], np.uint8), size=8192 * 40, p=[0.5, 0.5])   # long underflow runs
        assert np.array_equal(dev_encode(dev, data), O.encode(data))
        for name in ("m1m", "u1m_tail", "rr", "adv"):
            data = make_input(VECTORS[name])
            assert md5(dev_encode(dev, data)) == VECTORS[name]["payload_md5"]
    finally:
        _lib.set_option(_lib.OPT_ENCODE_PATH, _lib.ENCODE_AUTO)


def test_encode_kernels_agree_beyond_one_wave(dev):
    """The warp-specialised kernel also serves inputs of up to two resident waves of its CTAs (three per SM);
    125 MiB puts CTAs into a second wave.  The lane=packet kernel (oracle-checked above) is the yardstick."""
    from gpuar_b200 import _lib
    n = 8192 * 16000 + 123
    x = D.mixed_device(11, 126 << 20, 0)[:n]
    got = {}
    try:
        for name, path in (("fused", _lib.ENCODE_FUSED), ("ws", _lib.ENCODE_WS), ("auto", _lib.ENCODE_AUTO)):
            _lib.set_option(_lib.OPT_ENCODE_PATH, path)
            got[name] = dev.encode_bytes(x)
    finally:
        _lib.set_option(_lib.OPT_ENCODE_PATH, _lib.ENCODE_AUTO)
    assert torch.equal(got["fused"], got["ws"]) and torch.equal(got["fused"], got["auto"])
    assert torch.equal(dev.decode_bytes(got["ws"]), x)


@pytest.mark.parametrize("path", ["fused", "ws"])
@pytest.mark.parametrize("packet", [4096, 8192, 12288, 16112])
def test_packet_size_sweep_equals_rebuilt_reference(dev, packet, path):
    """BASELINE config 5: other packet sizes.  Goldens come from the reference rebuilt with gpu.h:12
    patched (tests/golden/sweep.json).  The device index/decoder must read them back."""
    from _vectors import SWEEP
    from gpuar_b200 import _lib
    rec = SWEEP[str(packet)]
    data = D.mixed(rec["seed"], rec["n"])
    _lib.set_option(_lib.OPT_ENCODE_PATH, _lib.ENCODE_FUSED if path == "fused" else _lib.ENCODE_WS)
    try:
        pay = dev.encode_bytes(to_dev(data), packet=packet)
    finally:
        _lib.set_option(_lib.OPT_ENCODE_PATH, _lib.ENCODE_AUTO)
    got = pay.cpu().numpy()
    assert got.size == rec["payload_bytes"] and md5(got) == rec["payload_md5"]
    assert np.array_equal(got, O.encode(data, packet))
    back = dev.decode_bytes(pay, packet=packet).cpu().numpy()
    assert np.array_equal(back, data)


@pytest.mark.parametrize("tile", [4, 8, 16, 32, 64, 128])
def test_compaction_work_unit_sweep(dev, tile):
    """BASELINE config 5, second half: 8192-byte packets, work unit of the scan + compaction kernel swept
    from 32 KiB to 1 MiB of input per CTA.  The stream must not depend on it."""
    from gpuar_b200 import _lib
    _lib.set_option(_lib.OPT_COMPACT_TILE, tile)
    try:
        for n in (1, 8192 * 3 + 7, 8192 * tile, 8192 * tile + 1, 8192 * (3 * tile + 1) + 100, 8192 * 700 + 33):
            data = D.mixed(n + tile, n)
            assert np.array_equal(dev_encode(dev, data), O.encode(data)), (tile, n)
        data = make_input(VECTORS["m1m"])
        assert md5(dev_encode(dev, data)) == VECTORS["m1m"]["payload_md5"]
    finally:
        _lib.set_option(_lib.OPT_COMPACT_TILE, 0)


# ------------------------------------------------------------------ index
@pytest.mark.parametrize("name", SMALL)
def test_index_equals_chain_walk(dev, name):
    data = make_input(VECTORS[name])
    pay = O.encode(data)
    c = pay.size
    padded = torch.zeros(c + 80, dtype=torch.uint8, device="cuda")
    padded[:c] = to_dev(pay)
    offsets, result = dev.index(padded, c, c // 5 + 1)
    packets, raw, status, _ = (int(v) for v in result.tolist())
    want = O.index(pay)
    assert status == 0 and packets == want.size and raw == data.size
    assert np.array_equal(offsets[:packets].cpu().numpy().astype(np.uint64), want)


def test_index_rejects_broken_chain(dev):
    pay = O.encode(D.uniform(3, 8192 * 5)).copy()
    pay[0] ^= 0x10                                   # first compLen now points into the middle of a packet
    c = pay.size
    padded = torch.zeros(c + 80, dtype=torch.uint8, device="cuda")
    padded[:c] = to_dev(pay)
    _, result = dev.index(padded, c, c // 5 + 1)
    assert int(result[2].item()) != 0


def test_index_small_packets_many_candidates(dev):
    # all-zero input: 210-byte packets, every one a candidate, ~40 per 8 KiB of payload
    data = D.zeros(8192 * 300)
    pay = O.encode(data)
    got = dev.decode_bytes(to_dev(pay)).cpu().numpy()
    assert np.array_equal(got, data)


# ------------------------------------------------------------------ decode
@pytest.mark.parametrize("name", SMALL)
def test_decode_reference_payload(dev, name):
    data = make_input(VECTORS[name])
    pay = O.encode(data)                             # produced by the CPU oracle, not by our encoder
    got = dev.decode_bytes(to_dev(pay)).cpu().numpy()
    assert np.array_equal(got, data)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 17, 8191, 8193, 8192 * 33 - 1, 8192 * 64 + 77])
def test_decode_ragged_sizes(dev, n):
    data = D.and3(n, n)
    got = dev.decode_bytes(to_dev(O.encode(data))).cpu().numpy()
    assert np.array_equal(got, data)


@pytest.mark.parametrize("path", ["latency", "throughput"])
def test_both_decode_kernels_are_bit_exact(dev, path):
    # the two decode kernels (decode_math.h: decode_step_latency / decode_step) on the same streams: golden
    # vectors made by the reference, ragged tails, a long-underflow stream, one-symbol packets
    from gpuar_b200 import _lib
    rng = np.random.default_rng(11)
    cases = [make_input(VECTORS[name]) for name in SMALL]
    cases += [D.mixed(5, 8192 * 70 + 123), D.zeros(8192 * 3 + 1), D.uniform(9, 8192 * 40 - 1),
              rng.choice(np.array([127, 128], np.uint8), size=8192 * 33), straddle_stream()]
    _lib.set_option(_lib.OPT_DECODE_PATH, {"latency": 1, "throughput": 2}[path])
    try:
        for data in cases:
            got = dev.decode_bytes(to_dev(O.encode(data))).cpu().numpy()
            assert np.array_equal(got, data)
    finally:
        _lib.set_option(_lib.OPT_DECODE_PATH, 0)


def test_random_distributions_all_kernels(dev):
    # one stream of packets with heavy skew, runs, ramps, two-symbol alphabets and sorted bytes (the fixed vectors
    # hold none of these) through every encode path and both decode kernels, against the oracle
    from gpuar_b200 import _lib
    rng = np.random.default_rng(20261017)
    parts = []
    for k in range(66):
        n = 8192
        kind = k % 6
        if kind == 0:
            parts.append(rng.choice(256, size=n, p=rng.dirichlet(np.full(256, 0.02))).astype(np.uint8))
        elif kind == 1:
            parts.append(np.repeat(rng.integers(0, 256, size=n // 37 + 1, dtype=np.uint8), 37)[:n])
        elif kind == 2:
            parts.append((np.arange(n) // int(rng.integers(1, 65))).astype(np.uint8))
        elif kind == 3:
            parts.append(rng.choice(np.array([0, 255], np.uint8), size=n, p=[0.97, 0.03]))
        elif kind == 4:
            parts.append(np.minimum(rng.geometric(0.3, size=n) - 1, 255).astype(np.uint8))
        else:
            parts.append(np.sort(rng.integers(0, 256, size=n, dtype=np.uint8)))
    data = np.concatenate(parts + [parts[0][:777]])
    want = O.encode(data)
    try:
        for enc in (0, 1, 2):
            _lib.set_option(_lib.OPT_ENCODE_PATH, enc)
            assert np.array_equal(dev_encode(dev, data), want), enc
        for dec in (1, 2):
            _lib.set_option(_lib.OPT_DECODE_PATH, dec)
            assert np.array_equal(dev.decode_bytes(to_dev(want)).cpu().numpy(), data), dec
    finally:
        _lib.set_option(_lib.OPT_ENCODE_PATH, 0)
        _lib.set_option(_lib.OPT_DECODE_PATH, 0)


def test_device_selfcheck_of_the_quotient():
    # divide_floor (one-sided float estimate) over every range and quotient, on the device
    from gpuar_b200 import _lib
    assert _lib.selfcheck() == 0


# ------------------------------------------------------------ host-buffer path
@pytest.mark.parametrize("name", ["one", "short", "maintest", "u8193", "m96k", "m1m", "u1m_tail"])
def test_gip_image_equals_reference_under_header_mask(codec, name):
    data = make_input(VECTORS[name])
    g = codec.compress(data)
    assert O.masked_equal(g, O.gip_file(data))
    assert int.from_bytes(g[12:16].tobytes(), "little") == g.size
    assert np.array_equal(codec.decompress(g), data)
    # and a reference-style image (garbage in the undefined header bytes) decodes too
    h = O.gip_file(data)
    for k in O.HEADER_MASKED:
        h[k] = 0xC3
    assert np.array_equal(codec.decompress(h, out_cap=data.size), data)


def test_host_path_multi_chunk(codec):
    # > 3 chunks of 2048 packets: exercises the rotating staging lanes of compress_host
    n = 8192 * 2048 * 4 + 12345
    data = D.mixed(21, n)
    g = codec.compress(data)
    ref = O.ref_encode(data, threads=8) if O.have_ref() else O.encode(data)
    assert np.array_equal(g[20:], ref)
    assert np.array_equal(codec.decompress(g), data)


def test_decompress_rejects_bad_magic(codec):
    from gpuar_b200._lib import GpuarError
    g = codec.compress(D.uniform(1, 9000)).copy()
    g[1] = 9
    with pytest.raises(GpuarError):
        codec.decompress(g, out_cap=9000)


# ------------------------------------------------------- reference-named shims
def test_shims_use_reference_slot_layout(codec):
    from gpuar_b200._lib import lib
    n = 8192 * 37 + 100
    data = D.uniform(77, n)
    packets = O.n_packets(n)
    src = torch.zeros(packets * 8192, dtype=torch.uint8, device="cuda")
    src[:n] = to_dev(data)
    slots = torch.zeros(packets * 8704, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    lib().initConstantRange()
    lib().garCompressExecutor(src.data_ptr(), n, slots.data_ptr(), 0)
    torch.cuda.synchronize()
    s = slots.cpu().numpy().reshape(packets, 8704)
    pay = O.encode(data)
    offs = O.index(pay)
    for p, o in enumerate(offs):
        ln = int(s[p, 0]) | (int(s[p, 1]) << 8)
        assert np.array_equal(s[p, :ln], pay[int(o): int(o) + ln])
    out = torch.zeros(packets * 8192, dtype=torch.uint8, device="cuda")
    lib().garDecompressExecutor(slots.data_ptr(), packets * 8704, out.data_ptr(), 0)
    torch.cuda.synchronize()
    assert np.array_equal(out[:n].cpu().numpy(), data)


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
def test_slots_equal_reference_gpu_kernel(codec):
    """The reference's own garCompress kernel (compiled from its sources for sm_100) on the same device."""
    from gpuar_b200._lib import lib
    n = 8192 * 64
    data = D.mixed(5, n)
    src = to_dev(data)
    ours = torch.zeros(64 * 8704, dtype=torch.uint8, device="cuda")
    theirs = torch.zeros(64 * 8704, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    O.ref().gpuar_ref_gpu_init()
    O.ref().gpuar_ref_gpu_encode(src.data_ptr(), n, theirs.data_ptr())
    assert O.ref().gpuar_ref_gpu_sync() == 0
    lib().garCompressExecutor(src.data_ptr(), n, ours.data_ptr(), 0)
    torch.cuda.synchronize()
    a, b = ours.cpu().numpy().reshape(64, 8704), theirs.cpu().numpy().reshape(64, 8704)
    for p in range(64):
        ln = int(b[p, 0]) | (int(b[p, 1]) << 8)
        assert np.array_equal(a[p, :ln], b[p, :ln])
    back = torch.zeros(n, dtype=torch.uint8, device="cuda")
    O.ref().gpuar_ref_gpu_decode(ours.data_ptr(), 64, back.data_ptr())
    assert O.ref().gpuar_ref_gpu_sync() == 0
    assert np.array_equal(back.cpu().numpy(), data)


# ------------------------------------------------------------ full-size cases
@pytest.mark.parametrize("name", LARGE)
def test_config2_64mib_md5(dev, name):
    """BASELINE config 2: 64 MiB uniform random, payload byte-identical to the reference (md5), md5 round trip."""
    rec = VECTORS[name]
    x = D.uniform_device(rec["seed"], rec["n"])
    assert hashlib.md5(x.cpu().numpy().tobytes()).hexdigest() == rec["input_md5"]
    pay = dev.encode_bytes(x)
    assert pay.numel() == rec["payload_bytes"]
    assert hashlib.md5(pay.cpu().numpy().tobytes()).hexdigest() == rec["payload_md5"]
    back = dev.decode_bytes(pay)
    assert torch.equal(back, x)


def test_config3_1gib_skewed_round_trip(dev):
    """BASELINE config 3: 1 GiB and3 (low entropy).  Size-independent properties: round trip, ratio,
    and prefix consistency (the first 1 MiB encodes to the golden s1m payload)."""
    n = 1 << 30
    x = D.and3_device(2, n)
    payload, total, _ = dev.encode(x)
    c = int(total.item())
    assert 0.55 < c / n < 0.57
    rec = VECTORS["s1m"]
    head = payload[: rec["payload_bytes"]].cpu().numpy()
    assert md5(head) == rec["payload_md5"]            # packets are independent: a prefix is a prefix
    packets = n // 8192
    offsets, result = dev.index(payload, c, packets)
    assert [int(v) for v in result.tolist()[:3]] == [packets, n, 0]
    back = dev.decode(payload, c, offsets, packets)
    assert torch.equal(back, x)


def test_beyond_4gib_header_and_round_trip(dev, codec):
    """sizes past 2^32.  The reference's header fields are 32 bit
    (file_header.hpp:61-72); ours carry the high halves in the bytes the reference leaves undefined."""
    free, _ = torch.cuda.mem_get_info()
    if free < 40 << 30:
        pytest.skip("needs ~40 GB of free device memory")
    n = (4 << 30) + (64 << 20) + 65536 * 3
    x = D.mixed_device(31, n)
    payload, total, _ = dev.encode(x)
    c = int(total.item())
    assert 0.5 * n < c < 0.7 * n
    packets = n // 8192
    offsets, result = dev.index(payload, c, packets)
    assert [int(v) for v in result.tolist()[:3]] == [packets, n, 0]
    assert int(offsets[packets - 1].item()) > (1 << 31)
    back = dev.decode(payload, c, offsets, packets)
    assert torch.equal(back, x)
    h = codec.write_header(n, 20 + c)
    assert int.from_bytes(h[4:12].tobytes(), "little") == n and int.from_bytes(h[4:8].tobytes(), "little") == n % (1 << 32)
    # prefix property: the first 1 MiB of the payload is the golden payload of mixed(31, 1 MiB)
    head = O.encode(D.mixed(31, 1 << 20))
    assert np.array_equal(payload[: head.size].cpu().numpy(), head)
    del back, offsets
    # SURVEY 8f-3 interop: the reference's CPU decoder (cpu_compressor.cpp:47-78 semantics: hops over the
    # compLen fields, ignores the header) reads the whole > 4 GiB device-encoded stream
    host_x = x.cpu().numpy()
    host_pay = payload[:c].cpu().numpy()
    if O.have_ref():
        got = O.ref_decode(host_pay, n, os.cpu_count() or 1)
        assert got.size == n and np.array_equal(got, host_x)
        del got
    # ... and the host-buffer path with an image of 4 GiB and more: written and read with 64-bit sizes
    # (marked in header byte 3), and read when the header is the reference's (32-bit field wrapped,
    # undefined bytes garbage): then the size comes from a hop over the packet headers
    del x, payload
    torch.cuda.empty_cache()
    g = codec.compress(host_x)
    assert g.size == 20 + c and g[3] == 0xB2 and codec.raw_size(g) == n
    assert np.array_equal(g[20:], host_pay)
    assert np.array_equal(codec.decompress(g), host_x)
    ref_style = g.copy()
    for k in O.HEADER_MASKED:
        ref_style[k] = 0x5A
    assert codec.raw_size(ref_style) == n % (1 << 32) and codec.walk(ref_style) == (packets, n)
    assert np.array_equal(codec.decompress(ref_style), host_x)


def test_corrupt_streams_never_fault(dev, codec):
    """Robustness the reference lacks (it may read out of bounds on corrupt input,
    gpuar_kernel.cu:746): damaged bitstreams decode to garbage of the right length, damaged chains are
    reported, and the device stays healthy."""
    from gpuar_b200._lib import GpuarError
    rng = np.random.default_rng(11)
    data = D.mixed(5, 8192 * 40 + 100)
    pay = O.encode(data)
    offs = O.index(pay).astype(np.int64)
    for trial in range(20):
        bad = pay.copy()
        hits = rng.integers(0, bad.size, size=8)
        hits = np.array([h for h in hits if not np.any((h >= offs) & (h < offs + 4))], dtype=np.int64)  # keep headers intact
        bad[hits] ^= rng.integers(1, 256, size=hits.size).astype(np.uint8)
        out = dev.decode_bytes(to_dev(bad)).cpu().numpy()
        assert out.size == data.size
    for trial in range(20):
        bad = pay.copy()
        o = int(offs[rng.integers(0, offs.size)])
        bad[o + rng.integers(0, 4)] ^= np.uint8(1 << rng.integers(0, 8))
        try:
            out = dev.decode_bytes(to_dev(bad)).cpu().numpy()
            assert out.size <= data.size + 8192
        except GpuarError:
            pass
    torch.cuda.synchronize()
    assert np.array_equal(dev.decode_bytes(to_dev(pay)).cpu().numpy(), data)      # the device is still fine


def test_abi_rejects_bad_arguments(dev, codec):
    from gpuar_b200 import _lib
    L = _lib.lib()
    x = torch.zeros(8192 * 2 + 16, dtype=torch.uint8, device="cuda")
    pay = torch.zeros(codec.payload_bound(8192 * 2) + 16, dtype=torch.uint8, device="cuda")
    tot = torch.zeros(1, dtype=torch.int64, device="cuda")
    scr = torch.zeros(int(L.gpuar_b200_encode_scratch_bytes(8192 * 2)) + 256, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ok = L.gpuar_b200_encode(x.data_ptr(), 8192 * 2, pay.data_ptr(), pay.numel(), tot.data_ptr(), None, scr.data_ptr(), scr.numel(), st)
    assert ok == 0
    # misaligned input, short payload buffer, short scratch, bad packet size
    assert L.gpuar_b200_encode(x.data_ptr() + 1, 8192, pay.data_ptr(), pay.numel(), tot.data_ptr(), None, scr.data_ptr(), scr.numel(), st) == _lib.E_ARG
    assert L.gpuar_b200_encode(x.data_ptr(), 8192 * 2, pay.data_ptr(), 100, tot.data_ptr(), None, scr.data_ptr(), scr.numel(), st) == _lib.E_ARG
    assert L.gpuar_b200_encode(x.data_ptr(), 8192 * 2, pay.data_ptr(), pay.numel(), tot.data_ptr(), None, scr.data_ptr(), 16, st) == _lib.E_ARG
    assert L.gpuar_b200_encode_ex(x.data_ptr(), 8192, 8200, pay.data_ptr(), pay.numel(), tot.data_ptr(), None, scr.data_ptr(), scr.numel(), st) == _lib.E_ARG
    assert L.gpuar_b200_payload_bound_ex(8192, 20000) == 0
    torch.cuda.synchronize()
