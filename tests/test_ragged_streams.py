"""Streams with short packets before the last one ("ragged" streams).

The reference never writes them (every packet but the last holds 8192 raw bytes,
cpu_compressor.cpp:144-173, gpu_compressor.cpp:100-131), but its CPU decoder accepts them: it
writes rawLen bytes per packet, back to back (cpu_compressor.cpp:60-70); its GPU decoder does not
(packet t is written at t*8192, gpuar_kernel.cu:924).  They arise when payloads are concatenated
(packets are self-delimiting).  Here: the device index flags them (result[3]) and
gpuar_b200_decode_packed / gpuar_b200_decompress_host / the CLI write them like the CPU decoder."""
import os
import subprocess

import numpy as np
import pytest

import _oracle as O
from gpuar_b200 import datagen as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "gpuar_b200", "gpuar")


def ragged_stream(lengths, seed=1):
    """Concatenated payloads of independently encoded pieces; returns (payload, plain text)."""
    pieces, pays = [], []
    for k, n in enumerate(lengths):
        gen = (D.uniform, D.and3, D.mixed)[k % 3]
        piece = gen(seed + k, n)
        pieces.append(piece)
        pays.append(O.encode(piece))
    return np.concatenate(pays), np.concatenate(pieces)


CASES = {
    "short_first": [100, 8192 * 3],
    "short_middle": [8192 * 2 + 5, 8192 * 4, 77],
    "every_piece_short": [8192 + 1, 8192 * 2 + 8191, 1, 4097, 8192 * 33 + 15],
    "short_then_full_tail": [8192 * 40 + 3000, 8192 * 64],
}


# ----------------------------------------------------------------- CPU: the oracle's semantics
@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_decodes_ragged_streams_like_the_reference_cpu_decoder(name):
    pay, plain = ragged_stream(CASES[name])
    assert np.array_equal(O.decode(pay, plain.size), plain)
    if O.have_ref():                                               # the reference's own arDecompress, packet by packet
        assert np.array_equal(O.ref_decode(pay, plain.size), plain)


# ----------------------------------------------------------------- GPU
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


def padded(pay, codec):
    buf = torch.zeros(pay.size + codec.PAD + 16, dtype=torch.uint8, device="cuda")
    buf[: pay.size] = to_dev(pay)
    return buf


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_device_index_flags_and_decode_packed_equals_oracle(dev, codec, name):
    pay, plain = ragged_stream(CASES[name])
    buf = padded(pay, codec)
    offsets, result = dev.index(buf, pay.size, pay.size // 5 + 1)
    packets, raw, status, ragged = (int(v) for v in result.tolist())
    want = O.index(pay)
    assert (status, ragged) == (0, 1)
    assert packets == want.size and raw == plain.size
    assert np.array_equal(offsets[:packets].cpu().numpy().astype(np.uint64), want)
    assert np.array_equal(dev.decode_bytes(to_dev(pay)).cpu().numpy(), plain)


@pytest.mark.gpu
def test_decode_packed_equals_decode_on_regular_streams(dev, codec):
    data = D.mixed(12, 8192 * 70 + 123)
    pay = O.encode(data)
    buf = padded(pay, codec)
    offsets, result = dev.index(buf, pay.size, pay.size // 5 + 1)
    packets, raw, status, ragged = (int(v) for v in result.tolist())
    assert (status, ragged, raw) == (0, 0, data.size)
    out = torch.zeros(raw + 16, dtype=torch.uint8, device="cuda")
    total = dev.decode_packed(buf, pay.size, offsets, packets, out)
    assert int(total.item()) == raw
    assert np.array_equal(out[:raw].cpu().numpy(), data)


@pytest.mark.gpu
def test_decode_packed_never_writes_past_the_capacity(dev, codec):
    pay, plain = ragged_stream(CASES["every_piece_short"])
    buf = padded(pay, codec)
    offsets, result = dev.index(buf, pay.size, pay.size // 5 + 1)
    packets = int(result[0].item())
    cap = 8192 * 5 + 16                                             # room for a few packets only
    out = torch.full((plain.size + 64,), 0xEE, dtype=torch.uint8, device="cuda")
    total = dev.decode_packed(buf, pay.size, offsets, packets, out[:cap])
    assert int(total.item()) == plain.size                          # the caller sees that it did not fit
    got = out.cpu().numpy()
    assert np.all(got[cap:] == 0xEE)
    # whole packets that fit are there
    offs = O.index(pay).astype(np.int64)
    raws = (pay[offs + 2].astype(np.int64) | (pay[offs + 3].astype(np.int64) << 8))
    ends = np.cumsum(raws)
    fit = int(ends[ends <= cap][-1])
    assert np.array_equal(got[:fit], plain[:fit])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_decompress_host_ragged(codec, name):
    pay, plain = ragged_stream(CASES[name])
    gip = np.concatenate([O.header(plain.size, 20 + pay.size), pay])
    assert np.array_equal(codec.decompress(gip), plain)


@pytest.mark.gpu
def test_decompress_host_many_short_packets_grow_the_offset_arrays(codec):
    """5000 packets of ~100 raw bytes: 60x more packets than a stream of full packets of that
    size has, so the offset arrays double several times while chunks are in flight."""
    rng = np.random.default_rng(3)
    lengths = rng.integers(1, 200, size=5000).tolist()
    pay, plain = ragged_stream(lengths, seed=50)
    gip = np.concatenate([O.header(plain.size, 20 + pay.size), pay])
    assert np.array_equal(codec.decompress(gip), plain)
    # a regular stream afterwards goes the usual way
    data = D.uniform(9, 8192 * 300 + 5)
    assert np.array_equal(codec.decompress(codec.compress(data)), data)


@pytest.mark.gpu
def test_decompress_host_ragged_across_chunks(codec):
    """Large enough that the host pipeline cuts several chunks (2 MiB of payload each), with short
    packets both inside chunks and at chunk ends."""
    lengths = [8192 * 300 + 17, 8192 * 255 + 1, 8192 * 256, 5, 8192 * 513 + 4000]
    pay, plain = ragged_stream(lengths, seed=7)
    gip = np.concatenate([O.header(plain.size, 20 + pay.size), pay])
    assert np.array_equal(codec.decompress(gip), plain)


@pytest.mark.gpu
def test_cli_decodes_concatenated_payloads(tmp_path):
    pay, plain = ragged_stream(CASES["short_then_full_tail"])
    gip = np.concatenate([O.header(plain.size, 20 + pay.size), pay])
    src, back = str(tmp_path / "cat.gip"), str(tmp_path / "back.dat")
    gip.tofile(src)
    for extra in ([], ["--segment=1"]):
        r = subprocess.run([CLI, "d", "--in", src, "--out", back, *extra], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        assert np.array_equal(np.fromfile(back, np.uint8), plain)
    # and the CPU mode agrees (the reference's CPU decoder semantics)
    r = subprocess.run([CLI, "d", "--host", "--in", src, "--out", back], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert np.array_equal(np.fromfile(back, np.uint8), plain)
