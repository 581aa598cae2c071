"""The `gpuar` command line (gpuar_b200/csrc/host/main.cpp): same flag surface as the reference
(src/main.cpp:83-107), .gip files interchangeable with the reference's."""
import os
import subprocess

import numpy as np
import pytest

import _oracle as O
from gpuar_b200 import datagen as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "gpuar_b200", "gpuar")


def run(*args):
    return subprocess.run([CLI, *args], capture_output=True, text=True, timeout=600)


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(CLI):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "gpuar_b200", "csrc"), "gpuar"],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def test_help_and_missing_input():
    assert "Usage: gpuar" in run("--help").stdout
    assert "Usage: gpuar" in run().stdout
    r = run("c", "--out=x.gip")
    assert r.returncode == 1 and "--in" in r.stderr


def test_host_mode_round_trip_and_reference_parity(tmp_path):
    data = D.mixed(3, (1 << 18) + 777)
    src, gip, back = (str(tmp_path / n) for n in ("in.dat", "out.gip", "back.dat"))
    data.tofile(src)
    r = run("c", "--host", "--in", src, f"--out={gip}")            # both argument spellings
    assert r.returncode == 0 and "Compression ratio" in r.stdout and "Attention: execute kernel code on host." in r.stdout
    g = np.fromfile(gip, np.uint8)
    assert O.masked_equal(g, O.gip_file(data))
    assert run("d", "--host", f"--in={gip}", "--out", back).returncode == 0
    assert np.array_equal(np.fromfile(back, np.uint8), data)


def test_host_mode_decodes_reference_style_image(tmp_path):
    data = D.and3(4, 30000)
    g = O.gip_file(data)
    for k in O.HEADER_MASKED:
        g[k] = 0x5A                                                # garbage the reference leaves there
    gip, back = str(tmp_path / "ref.gip"), str(tmp_path / "back.dat")
    g.tofile(gip)
    assert run("d", "--host", "--in", gip, "--out", back).returncode == 0
    assert np.array_equal(np.fromfile(back, np.uint8), data)


def test_empty_file(tmp_path):
    src, gip, back = (str(tmp_path / n) for n in ("e.dat", "e.gip", "e.back"))
    open(src, "wb").close()
    assert run("c", "--host", "--in", src, "--out", gip).returncode == 0
    assert os.path.getsize(gip) == 20
    assert run("d", "--host", "--in", gip, "--out", back).returncode == 0
    assert os.path.getsize(back) == 0


def test_no_device_is_an_error_not_a_cpu_fallback(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    src = str(tmp_path / "in.dat")
    D.uniform(1, 9000).tofile(src)
    r = run("c", "--in", src, "--out", str(tmp_path / "o.gip"))
    assert r.returncode == 1 and "no CUDA device" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("segment_mib", [1024, 1])
def test_gpu_mode_files_equal_reference(tmp_path, segment_mib):
    n = (5 << 20) + 4321                                            # several 1 MiB segments and a ragged tail
    data = D.mixed(8, n)
    src, gip, back = (str(tmp_path / n_) for n_ in ("in.dat", "out.gip", "back.dat"))
    data.tofile(src)
    r = run("c", "--in", src, "--out", gip, f"--segment={segment_mib}")
    assert r.returncode == 0, r.stderr
    g = np.fromfile(gip, np.uint8)
    assert O.masked_equal(g, O.gip_file(data))
    r = run("d", "--in", gip, "--out", back, f"--segment={segment_mib}")
    assert r.returncode == 0, r.stderr
    assert np.array_equal(np.fromfile(back, np.uint8), data)
    # and the CPU mode reads what the GPU mode wrote
    assert run("d", "--host", "--in", gip, "--out", back).returncode == 0
    assert np.array_equal(np.fromfile(back, np.uint8), data)


@pytest.mark.gpu
@pytest.mark.parametrize("announced", [None, 0, 8192, 0xFFFFFFFF])
def test_gpu_mode_decode_sizes_its_staging_without_trusting_the_header(tmp_path, announced):
    """GpuCompressor::decompress sizes its page-locked staging from the announced raw size; the field
    is 32 bits in the reference (file_header.hpp:61-66) and wraps for big files, so a wrong value
    may only change how the stream is cut into segments, never the result."""
    data = D.and3(5, (3 << 20) + 100)                               # payload ~0.56 of the raw size
    g = O.gip_file(data)
    for k in O.HEADER_MASKED:
        g[k] = 0xA5
    if announced is not None:
        g[4:8] = np.frombuffer(int(announced).to_bytes(4, "little"), np.uint8)
    gip, back = str(tmp_path / "in.gip"), str(tmp_path / "back.dat")
    g.tofile(gip)
    r = run("d", "--in", gip, "--out", back)
    assert r.returncode == 0, r.stderr
    assert np.array_equal(np.fromfile(back, np.uint8), data)
