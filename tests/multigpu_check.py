#!/usr/bin/env python
"""Multi-GPU parity check (not a pytest file: needs W GPUs and torchrun).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multigpu_check.py

Every rank encodes its packet range of one input; the streams are concatenated into rank 0's
buffer by gpuar_b200_shard_concat over NVLink; rank 0 checks that the gathered payload is
byte-identical to the oracle's payload of the whole input (= the single-GPU payload), then
decodes it on its own GPU and compares with the input.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import _oracle as O  # noqa: E402
from gpuar_b200 import codec, datagen as D  # noqa: E402
from gpuar_b200.shard import ShardedCodec, byte_range  # noqa: E402


def main():
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = codec.DeviceCodec(local)
    ok = True
    for layout in ("segments", "gather"):
        sh = ShardedCodec(dev, rank, world, layout)
        for n in (8192 * 64 * world + 4321, 8192 * 3, 12 << 20):
            data = D.mixed(17, n)
            b0, b1 = byte_range(n, rank, world)
            x = torch.from_numpy(data[b0:b1].copy()).cuda()
            cap = codec.payload_bound(max(b1 - b0, 8192))
            sh.reserve(codec.payload_bound(n))
            payload, total, _ = dev.encode(x) if b1 > b0 else (torch.zeros(cap + 16, dtype=torch.uint8, device="cuda"),
                                                               torch.zeros(1, dtype=torch.int64, device="cuda"), None)
            sh.concat(payload, total)
            torch.cuda.synchronize()
            dist.barrier()
            # every owner hands its segment to rank 0 (gloo-style gather through NCCL, test only)
            seg, valid = sh._peer.my_segment()
            sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
            dist.all_gather(sizes, torch.tensor([valid], dtype=torch.int64, device="cuda"))
            sizes = [int(v.item()) for v in sizes]
            pieces = []
            for g in range(world):
                if sizes[g] == 0:
                    continue
                if g == 0:
                    if rank == 0:
                        pieces.append(seg.clone())
                elif rank == g:
                    dist.send(seg.contiguous(), dst=0)
                elif rank == 0:
                    buf = torch.empty(sizes[g], dtype=torch.uint8, device="cuda")
                    dist.recv(buf, src=g)
                    pieces.append(buf)
            if rank == 0:
                want = O.ref_encode(data, 8) if O.have_ref() else O.encode(data)
                got = torch.cat(pieces).cpu().numpy() if pieces else np.zeros(0, np.uint8)
                same = got.size == want.size and np.array_equal(got, want)
                back = dev.decode_bytes(torch.from_numpy(got).cuda()).cpu().numpy() if same else np.zeros(0, np.uint8)
                rt = np.array_equal(back, data)
                print(f"layout={layout} n={n} world={world} segment bytes={sizes} concatenated==single-GPU payload: "
                      f"{same}; round trip: {rt}", flush=True)
                ok = ok and same and rt
            dist.barrier()
        sh._peer.release()
    if rank == 0:
        print("MULTIGPU PARITY OK" if ok else "MULTIGPU PARITY FAILED", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
