#!/usr/bin/env python
"""Regenerate tests/golden/ from the REFERENCE's own codec.

Run in the build container (needs /root/reference and nvcc):

    python tests/golden/make_golden.py

Every payload below is produced by the reference's arCompress (gpuar_kernel.cu:487)
through oracle/_ref/libgpuar_ref.so (built by `make -C oracle ref` from the sources
in place) and round-tripped through the reference's arDecompress (:848).  Outputs:

  vectors.json       name -> {gen, n, input_md5, payload_bytes, payload_md5, packets}
  *.in / *.payload   small fixtures stored verbatim (inputs that have no generator,
                     payloads short enough to diff by eye)
  sweep.json         packet-size sweep {4096, 8192, 12288, 16112} on mixed(3, 1 MiB): the
                     reference rebuilt from a scratch copy under /tmp with gpu.h:12
                     patched (the repo never holds reference sources)

`maintest` is the 705-byte literal of the reference's dormant self-test
(main.cpp:9); it is extracted from the reference at generation time.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import _oracle as O  # noqa: E402
from gpuar_b200 import datagen as D  # noqa: E402

REF = os.environ.get("GPUAR_REFERENCE", "/root/reference")


def md5(a) -> str:
    return hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()


def maintest_literal() -> np.ndarray:
    src = open(os.path.join(REF, "src", "main.cpp"), encoding="utf-8", errors="replace").read()
    m = re.search(r'char original\[\] = "([^"]*)";', src)
    assert m, "self-test literal not found in the reference's main.cpp"
    return np.frombuffer(m.group(1).encode("ascii"), dtype=np.uint8)


def cases():
    return [
        # name, recipe (re-creatable by tests without fixtures), data
        ("one", {"gen": "bytes", "hex": "00"}, np.zeros(1, np.uint8)),
        ("short", {"gen": "bytes", "hex": b"abracadabra".hex()}, np.frombuffer(b"abracadabra", np.uint8)),
        ("maintest", {"gen": "file", "file": "maintest.in"}, maintest_literal()),
        ("zeros", {"gen": "zeros", "n": 1 << 20}, D.zeros(1 << 20)),
        ("rr", {"gen": "round_robin", "n": 1 << 20}, D.round_robin(1 << 20)),
        ("adv", {"gen": "adversarial_x4"}, np.tile(D.adversarial_packet(), 4)),
        ("u8193", {"gen": "uniform", "seed": 7, "n": 8193}, D.uniform(7, 8193)),
        ("u8191", {"gen": "uniform", "seed": 8, "n": 8191}, D.uniform(8, 8191)),
        ("u16384", {"gen": "uniform", "seed": 11, "n": 16384}, D.uniform(11, 16384)),
        ("m96k", {"gen": "mixed", "seed": 9, "n": 98304}, D.mixed(9, 98304)),
        ("u1m", {"gen": "uniform", "seed": 1, "n": 1 << 20}, D.uniform(1, 1 << 20)),
        ("s1m", {"gen": "and3", "seed": 2, "n": 1 << 20}, D.and3(2, 1 << 20)),
        ("m1m", {"gen": "mixed", "seed": 3, "n": 1 << 20}, D.mixed(3, 1 << 20)),
        ("u1m_tail", {"gen": "uniform", "seed": 5, "n": (1 << 20) + 4321}, D.uniform(5, (1 << 20) + 4321)),
        ("u64m", {"gen": "uniform", "seed": 0x64, "n": 64 << 20}, D.uniform(0x64, 64 << 20)),
    ]


def build_sweep_ref(packet: int, tmp: str) -> str:
    """Reference rebuilt with UNCOMPRESSED_PACKET_SIZE = packet, from a copy under /tmp."""
    dst = os.path.join(tmp, f"ref_{packet}")
    shutil.copytree(REF, dst)
    subprocess.check_call(["chmod", "-R", "u+w", dst])
    gpu_h = os.path.join(dst, "src", "gpu.h")
    s = open(gpu_h).read()
    s2 = s.replace("(8192 + EXTRA_COMPRESSED_SIZE)", f"({packet} + EXTRA_COMPRESSED_SIZE)")
    assert s2 != s or packet == 8192
    open(gpu_h, "w").write(s2)
    so = os.path.join(tmp, f"libref_{packet}.so")
    subprocess.check_call([
        "nvcc", "-O3", "--std=c++14", "-include", "cstdint", "-diag-suppress", "20040",
        "-Wno-deprecated-gpu-targets", "-Xcompiler", "-fPIC", "-shared", "-Xlinker", "-Bsymbolic",
        f"-I{dst}/src", f"-I{dst}/common", "-gencode", "arch=compute_100,code=sm_100",
        f"{dst}/src/gpuar_kernel.cu", os.path.join(ROOT, "oracle", "ref_harness.cu"), "-o", so])
    return so


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    threads = os.cpu_count() or 1
    vectors = {}
    for name, recipe, data in cases():
        pay = O.ref_encode(data, threads)
        back = O.ref_decode(pay, data.size, threads)
        assert np.array_equal(back, data), name
        offs = O.index(pay)
        vectors[name] = dict(recipe, n=int(data.size), input_md5=md5(data), payload_bytes=int(pay.size),
                             payload_md5=md5(pay), packets=int(offs.size))
        if recipe["gen"] == "file":
            data.tofile(os.path.join(HERE, recipe["file"]))
        if pay.size <= 16384:
            pay.tofile(os.path.join(HERE, f"{name}.payload"))
        print(f"{name:10s} n={data.size:9d} payload={pay.size:9d} {md5(pay)}")
    json.dump(vectors, open(os.path.join(HERE, "vectors.json"), "w"), indent=1, sort_keys=True)

    sweep = {}
    data = D.mixed(3, 1 << 20)
    with tempfile.TemporaryDirectory() as tmp:
        for packet in (4096, 8192, 12288, 16112):
            lib = C.CDLL(build_sweep_ref(packet, tmp))
            assert lib.gpuar_ref_packet_bytes() == packet
            lib.gpuar_ref_encode_stream.restype = C.c_size_t
            lib.gpuar_ref_encode_stream.argtypes = [O._u8p, C.c_size_t, O._u8p, C.c_int]
            buf = np.zeros(O.n_packets(data.size, packet) * (packet + 512) + 16, np.uint8)
            c = lib.gpuar_ref_encode_stream(O._ptr(data), data.size, O._ptr(buf), threads)
            pay = buf[:c]
            sweep[str(packet)] = {"gen": "mixed", "seed": 3, "n": 1 << 20, "payload_bytes": int(c),
                                  "payload_md5": md5(pay), "packets": O.n_packets(data.size, packet)}
            print(f"sweep {packet:6d} payload={c:8d} {md5(pay)}")
    json.dump(sweep, open(os.path.join(HERE, "sweep.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
