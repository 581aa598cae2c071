#!/usr/bin/env python
"""One profiled pass of the hot path for ncu (not a bench line): warm-up, then cudaProfilerStart .. one encode,
one chain discovery, one decode .. cudaProfilerStop, so that `ncu --profile-from-start off` sees exactly one
launch of every kernel of a step.

    ncu --set full --clock-control none --import-source on --profile-from-start off -o OUT python tools/ncu_target.py --workload s1g
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gpuar_b200 import codec  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="u64m", choices=sorted(bench.WORKLOADS))
    args = ap.parse_args()
    gen, seed, nbytes, _ = bench.WORKLOADS[args.workload]
    dev = codec.DeviceCodec(0)
    x = bench.gen_device(gen, seed, 0, nbytes)
    packets = (nbytes + 8191) // 8192
    payload = torch.empty(codec.payload_bound(nbytes) + 16, dtype=torch.uint8, device="cuda")
    total = torch.zeros(1, dtype=torch.int64, device="cuda")
    offsets = torch.empty(packets + 1, dtype=torch.int64, device="cuda")
    result = torch.zeros(4, dtype=torch.int64, device="cuda")
    out = torch.empty(packets * 8192, dtype=torch.uint8, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def step():
        dev.encode(x, payload, total)
        c = int(total.item())
        dev.index(payload, c, packets, offsets, result)
        dev.decode(payload, c, offsets, packets, out)
        return c

    for _ in range(2):
        c = step()
    torch.cuda.synchronize()
    assert torch.equal(out[:nbytes], x)
    flush.fill_(1)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print(f"{args.workload}: {nbytes} bytes -> {c} payload bytes", flush=True)


if __name__ == "__main__":
    main()
