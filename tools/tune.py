#!/usr/bin/env python
"""Kernel-level tuning sweep on one GPU (not a bench line): CUDA-event spans of the library's own
profiling hooks for encode / compaction / index / decode over input sizes, encoder paths and
compaction work units.  One process, so the numbers of one run compare with each other; select a
tuning variant of the library with GPUAR_B200_LIB (tools/build_variant.sh).

    python tools/tune.py [--sizes 64,128,256] [--paths ws,fused] [--tiles 0,8,64] [--reps 5]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpuar_b200 import _lib, codec, datagen as D  # noqa: E402


def spans(fn, reps, flush):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    _lib.profile(True)
    _lib.profile_read()
    for _ in range(reps):
        flush.fill_(1)                        # 256 MiB write: evicts the 126 MB L2
        fn()
    torch.cuda.synchronize()
    out = {k: round(ms / max(n, 1), 4) for k, (ms, n) in _lib.profile_read().items() if n}
    _lib.profile(False)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="64")
    ap.add_argument("--paths", default="auto")
    ap.add_argument("--tiles", default="0")
    ap.add_argument("--gen", default="uniform", choices=["uniform", "and3", "mixed"])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--decode", action="store_true")
    ap.add_argument("--dec-paths", default="0", help="GPUAR_OPT_DECODE_PATH values to time (0 auto, 1 latency, 2 throughput)")
    args = ap.parse_args()
    dev = codec.DeviceCodec(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    gen = {"uniform": D.uniform_device, "and3": D.and3_device, "mixed": D.mixed_device}[args.gen]
    for mib in (int(v) for v in args.sizes.split(",")):
        n = mib << 20
        x = gen(7, n, 0)
        payload = torch.empty(codec.payload_bound(n) + 16, dtype=torch.uint8, device="cuda")
        total = torch.zeros(1, dtype=torch.int64, device="cuda")
        for path in args.paths.split(","):
            _lib.set_option(_lib.OPT_ENCODE_PATH, {"auto": 0, "fused": 1, "ws": 2}[path])
            for tile in (int(v) for v in args.tiles.split(",")):
                if tile or _lib.lib().gpuar_b200_set_option(_lib.OPT_COMPACT_TILE, 0) == 0:
                    _lib.set_option(_lib.OPT_COMPACT_TILE, tile)          # (older variants lack the option)
                rec = spans(lambda: dev.encode(x, payload, total), args.reps, flush)
                line = {"lib": os.path.basename(_lib.lib()._name), "gen": args.gen, "mib": mib, "path": path,
                        "tile": tile, **rec}
                if args.decode:
                    c = int(total.item())
                    packets = (n + 8191) // 8192
                    off, res = dev.index(payload, c, packets)
                    out = torch.empty(packets * 8192, dtype=torch.uint8, device="cuda")
                    for dp in (int(v) for v in args.dec_paths.split(",")):
                        _lib.set_option(_lib.OPT_DECODE_PATH, dp)
                        out.zero_()
                        rec = spans(lambda: (dev.index(payload, c, packets, off, res),
                                             dev.decode(payload, c, off, packets, out)), args.reps, flush)
                        assert torch.equal(out[:n], x)
                        line.update(rec if dp == 0 else {f"decode_path{dp}": rec["decode"]})
                    _lib.set_option(_lib.OPT_DECODE_PATH, 0)
                print(json.dumps(line), flush=True)
        _lib.set_option(_lib.OPT_ENCODE_PATH, 0)
        _lib.lib().gpuar_b200_set_option(_lib.OPT_COMPACT_TILE, 0)
        del x, payload


if __name__ == "__main__":
    main()
