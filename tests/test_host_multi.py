"""gpuar_b200_compress_host_multi / decompress_host_multi: one process, one host thread, the chunks of one
input rotating over several GPUs (single-process form; the reference has one device only,
gpu_compressor.cpp:67-82).  The image must not depend on how many devices wrote it."""
import os
import subprocess

import numpy as np
import pytest

import _oracle as O
from gpuar_b200 import datagen as D

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "gpuar_b200", "gpuar")


@pytest.fixture(scope="module")
def codec():
    from gpuar_b200 import codec as cd
    cd.init()
    return cd


def device_sets():
    n = torch.cuda.device_count()
    sets = [[0]]
    if n >= 2:
        sets += [[0, 1], [1, 0]]
    if n >= 3:
        sets.append(list(range(n)))
    return sets


@pytest.mark.parametrize("n", [0, 1, 5000, 8192 * 300 + 17, (40 << 20) + 8192 * 3 + 5])
def test_image_is_independent_of_the_devices(codec, n):
    data = D.mixed(21, n)
    want = O.gip_file(data)
    for devs in device_sets():
        g = codec.compress(data, devices=devs)
        assert O.masked_equal(g, want), f"devices {devs}"
        assert np.array_equal(codec.decompress(g, devices=devs), data), f"devices {devs}"


def test_ragged_stream_over_several_devices(codec):
    """short packets in the middle of the stream (legal for the reference's CPU decoder, cpu_compressor.cpp:60-70)"""
    parts = [D.uniform(3, 8192 * 400 + 77), D.and3(4, 8192 * 300 + 1), D.mixed(5, (8 << 20) + 4000)]
    pay = np.concatenate([O.encode(p) for p in parts])
    raw = np.concatenate(parts)
    g = np.concatenate([O.header(raw.size, 20 + pay.size), pay])
    for devs in device_sets():
        assert np.array_equal(codec.decompress(g, devices=devs), raw), f"devices {devs}"


def test_bad_device_lists_are_rejected(codec):
    from gpuar_b200._lib import GpuarError
    data = D.uniform(1, 8192)
    for devs in ([0, 0], [torch.cuda.device_count()], []):
        with pytest.raises(GpuarError):
            codec.compress(data, devices=devs)


def test_cli_gpus_flag(tmp_path):
    n_dev = torch.cuda.device_count()
    data = D.mixed(9, (24 << 20) + 1234)
    src, gip, back = (str(tmp_path / f) for f in ("in.dat", "out.gip", "back.dat"))
    data.tofile(src)
    for gpus in sorted({1, min(2, n_dev), n_dev}):
        r = subprocess.run([CLI, "c", f"--gpus={gpus}", f"--in={src}", f"--out={gip}"], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        assert O.masked_equal(np.fromfile(gip, np.uint8), O.gip_file(data))
        r = subprocess.run([CLI, "d", f"--gpus={gpus}", f"--in={gip}", f"--out={back}"], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        assert np.array_equal(np.fromfile(back, np.uint8), data)
    r = subprocess.run([CLI, "c", f"--gpus={n_dev + 1}", f"--in={src}", f"--out={gip}"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 1
