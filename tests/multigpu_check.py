#!/usr/bin/env python
"""Multi-GPU parity check (not a pytest file: needs W GPUs and torchrun).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multigpu_check.py

Every rank encodes its packet range of one input with gpuar_b200_encode_sharded (totals through
the peer-mapped mailboxes, packets written straight to their place in the concatenated stream
over NVLink).  Checked, for both layouts and for inputs with ragged tails and with fewer packets
than ranks:
  * the concatenated stream is byte-identical to the reference's payload of the whole input
    (oracle/_ref when built, else the oracle port) = the single-GPU payload;
  * rank 0 decodes the gathered stream on its own GPU: equals the input;
  * "segments" layout: gpuar_b200_decode_sharded -- every rank discovers the chain of its own
    segment and decodes the packets that start there -- yields exactly the input, slice by slice.
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


def collect_on_rank0(view, valid, rank, world):
    """Every owner hands its part of the stream to rank 0 (through NCCL: test plumbing only)."""
    sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([valid], dtype=torch.int64, device="cuda"))
    sizes = [int(v.item()) for v in sizes]
    pieces = []
    for g in range(world):
        if sizes[g] == 0:
            continue
        if g == 0:
            if rank == 0:
                pieces.append(view.clone())
        elif rank == g:
            dist.send(view.contiguous(), dst=0)
        elif rank == 0:
            buf = torch.empty(sizes[g], dtype=torch.uint8, device="cuda")
            dist.recv(buf, src=g)
            pieces.append(buf)
    return pieces, sizes


def main():
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = codec.DeviceCodec(local)
    ok = True
    sizes_to_try = (8192 * 64 * world + 4321, 8192 * 3, 12 << 20, (40 << 20) + 8192 * 5 + 1, 5000)
    for layout in ("segments", "gather"):
        sh = ShardedCodec(dev, rank, world, layout)
        for rep, n in enumerate(sizes_to_try):
            data = D.mixed(17, n) if rep != 3 else D.uniform(23, n)
            b0, b1 = byte_range(n, rank, world)
            x = torch.from_numpy(data[b0:b1].copy()).cuda()
            sh.reserve(codec.payload_bound(max(b1 - b0, 8192)))
            for _ in range(3):                               # repeated calls: the mailbox tags and parities move on
                sh.encode(x)
            torch.cuda.synchronize()
            dist.barrier()
            view, valid = sh.group.my_segment()
            total = int(sh.group.layout_out[0].item())
            pieces, sizes = collect_on_rank0(view, valid, rank, world)
            good = True
            if rank == 0:
                want = O.ref_encode(data, 8) if O.have_ref() else O.encode(data)
                got = torch.cat(pieces).cpu().numpy() if pieces else np.zeros(0, np.uint8)
                same = got.size == want.size == total and np.array_equal(got, want)
                back = dev.decode_bytes(torch.from_numpy(got).cuda()).cpu().numpy() if same else np.zeros(0, np.uint8)
                rt = np.array_equal(back, data)
                print(f"layout={layout} n={n} world={world} segment bytes={sizes} concatenated==single-GPU payload: "
                      f"{same}; round trip: {rt}", flush=True)
                good = same and rt
            if layout == "segments":
                packets = (n + 8191) // 8192
                out = torch.zeros((packets + 2) * 8192, dtype=torch.uint8, device="cuda")
                result = torch.zeros(8, dtype=torch.int64, device="cuda")
                for _ in range(2):
                    sh.decode(total, out, result)
                torch.cuda.synchronize()
                mine, raw, status, before, raw_before = (int(v) for v in result.tolist()[:5])
                mine_ok = status == 0 and raw_before == min(before * 8192, n) and \
                    np.array_equal(out[:raw].cpu().numpy(), data[raw_before: raw_before + raw])
                counts = [torch.zeros(2, dtype=torch.int64, device="cuda") for _ in range(world)]
                dist.all_gather(counts, torch.tensor([mine, raw], dtype=torch.int64, device="cuda"))
                covered = sum(int(c[1].item()) for c in counts) == n and sum(int(c[0].item()) for c in counts) == packets
                flag = torch.tensor([1 if (mine_ok and covered) else 0], device="cuda")
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                if rank == 0:
                    print(f"    sharded decode: packets per rank {[int(c[0].item()) for c in counts]}, "
                          f"every rank's slice == input: {bool(flag.item())}", flush=True)
                good = good and bool(flag.item())
            ok = ok and good
            dist.barrier()
        sh.group.release()
    if rank == 0:
        print("MULTIGPU PARITY OK" if ok else "MULTIGPU PARITY FAILED", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
