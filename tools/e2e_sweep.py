#!/usr/bin/env python
"""Host-buffer pipeline timing for one chunk count (GPUAR_B200_HOST_CHUNKS, read once per process)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpuar_b200 import codec, datagen as D
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = mib << 20
x = D.uniform_device(0x64, n, 0)
hin = torch.empty(n, dtype=torch.uint8).pin_memory(); hin.copy_(x)
hg = torch.empty(20 + codec.payload_bound(n), dtype=torch.uint8).pin_memory()
ho = torch.empty(n + 8192, dtype=torch.uint8).pin_memory()
a, g_, o = hin.numpy(), hg.numpy(), ho.numpy()
g = codec.compress(a, out=g_); codec.decompress(g, out=o)
best = [1e9, 1e9]
for _ in range(8):
    t = time.perf_counter(); g = codec.compress(a, out=g_); best[0] = min(best[0], time.perf_counter() - t)
    t = time.perf_counter(); b = codec.decompress(g, out=o); best[1] = min(best[1], time.perf_counter() - t)
assert np.array_equal(b[:n], a)
print(f"chunks={os.environ.get('GPUAR_B200_HOST_CHUNKS','dflt'):>4s} {mib} MiB: compress {best[0]*1e3:.3f} ms = {n/best[0]/1e9:.1f} GB/s, decompress {best[1]*1e3:.3f} ms = {n/best[1]/1e9:.1f} GB/s")
