#!/usr/bin/env python
"""Where the step time of the sharded encode goes at N > 1 (torchrun; not a bench line): per-rank CUDA-event
times of the local encode, the sharded encode, and its two kernel groups, with and without the L2 flush.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tools/shard_probe.py [--mib 64]
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpuar_b200 import _lib, codec, datagen as D  # noqa: E402
from gpuar_b200.shard import ShardedCodec  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = codec.DeviceCodec(local)
    n = args.mib << 20
    x = D.uniform_device(0x64, n, rank * n)
    sh = ShardedCodec(dev, rank, world, "segments")
    sh.reserve(codec.payload_bound(n))
    payload = torch.empty(codec.payload_bound(n) + 16, dtype=torch.uint8, device="cuda")
    total = torch.zeros(1, dtype=torch.int64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def timed(step, k, do_flush):
        evs = []
        for _ in range(k):
            if do_flush:
                flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); step(); b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        t = sorted(a.elapsed_time(b) for a, b in evs)
        return {"mean": sum(t) / len(t), "min": t[0], "med": t[len(t) // 2], "max": t[-1]}

    waits = []

    def sharded_step():
        sh.encode(x)
        waits.append(sh.group.layout_out.clone())

    out = {}
    for name, step in (("local", lambda: dev.encode(x, payload, total)), ("sharded", sharded_step)):
        for _ in range(5):
            step()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        for fl in (True, False):
            _lib.profile(True); _lib.profile_read()
            out[f"{name}{'_flush' if fl else ''}"] = timed(step, args.steps, fl)
            sp = _lib.profile_read(); _lib.profile(False)
            out[f"{name}{'_flush' if fl else ''}_spans"] = {k: round(v[0] / max(1, v[1]), 4) for k, v in sp.items() if v[1]}
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    w = sorted(int(t[5].item()) / 1e3 for t in waits[-2 * args.steps:])
    out["totals_wait_us"] = {"min": w[0], "med": w[len(w) // 2], "max": w[-1]}
    # tracing build (tools/build_variant.sh trace -DGPUAR_SHARD_TRACE): per-tile time line of the last compaction
    try:
        import ctypes as C
        import numpy as np
        fn = _lib.lib().gpuar_b200_debug_shard_trace
        tiles = (n // 8192 + 7) // 8
        tr = np.zeros(tiles * 4, dtype=np.uint64)
        torch.cuda.synchronize()
        assert fn(tr.ctypes.data_as(C.c_void_p), C.c_size_t(tr.size)) == 0
        tr = tr.reshape(tiles, 4).astype(np.int64)
        t0 = int(tr[:, 0].min())
        q = [0, tiles // 4, tiles // 2, 3 * tiles // 4, tiles - 1]
        out["trace_us"] = {"tiles": tiles, "first_start": 0.0, "last_start": round((int(tr[:, 0].max()) - t0) / 1e3, 1),
                           "last_end": round((int(tr[:, 3].max()) - t0) / 1e3, 1),
                           "lookback_done_max": round((int(tr[:, 1].max()) - t0) / 1e3, 1),
                           "totals_done_max": round((int(tr[:, 2].max()) - t0) / 1e3, 1),
                           "sample_tiles[start,lookback,totals,end]": {int(t): [round((int(v) - t0) / 1e3, 1) for v in tr[t]] for t in q},
                           "copy_us_mean": round(float((tr[:, 3] - tr[:, 2]).mean()) / 1e3, 1),
                           "copy_us_max": round(float((tr[:, 3] - tr[:, 2]).max()) / 1e3, 1)}
    except AttributeError:
        pass
    rows = [None] * world
    dist.all_gather_object(rows, out)
    if rank == 0:
        for r, row in enumerate(rows):
            print(json.dumps({"rank": r, **{k: (round(v["mean"], 4) if "mean" in v else v) for k, v in row.items() if k != "stamps"}}))
        for key in ("local_flush", "sharded_flush", "local", "sharded"):
            print(key, "max over ranks of mean ms:", round(max(r[key]["mean"] for r in rows), 4),
                  "min/med/max of rank 0:", {k: round(v, 4) for k, v in rows[0][key].items()})
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
