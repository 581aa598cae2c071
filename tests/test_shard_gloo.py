"""Host-side sharding logic on CPU with torch.distributed (gloo, world_size 2 and 3): packet
ranges partition the input, the exclusive scan of the per-rank payload totals gives landing
offsets, and the concatenation equals the payload of the whole input.  The per-rank encoder
here is the oracle (test infrastructure); on the GPU box the same plumbing drives the CUDA path
(tests/multigpu_check.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import _oracle as O
from gpuar_b200 import datagen as D
from gpuar_b200.shard import byte_range, exclusive_scan, packet_range


@pytest.mark.parametrize("n", [0, 1, 8191, 8192, 8193, 8192 * 7 + 5, 1 << 20])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_ranges_partition_the_input(n, world):
    pieces = [byte_range(n, r, world) for r in range(world)]
    assert pieces[0][0] == 0 and pieces[-1][1] == n
    for (a0, a1), (b0, b1) in zip(pieces, pieces[1:]):
        assert a1 == b0 and a0 <= a1
    for r in range(world):
        p0, p1 = packet_range(n, r, world)
        assert pieces[r][0] == min(n, p0 * 8192)            # shards start on packet boundaries
    sizes = [p1 - p0 for p0, p1 in (packet_range(n, r, world) for r in range(world))]
    assert max(sizes) - min(sizes) <= 1


def test_exclusive_scan():
    assert exclusive_scan([5, 0, 7]) == [0, 5, 5]
    assert exclusive_scan([]) == []


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data = D.mixed(11, n)
    b0, b1 = byte_range(n, rank, world)
    mine = O.encode(data[b0:b1])
    totals = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(totals, torch.tensor([mine.size], dtype=torch.int64))
    totals = [int(t.item()) for t in totals]
    offs = exclusive_scan(totals)
    if rank == 0:
        out = np.zeros(sum(totals), np.uint8)
        out[: mine.size] = mine
        for r in range(1, world):
            buf = torch.zeros(totals[r], dtype=torch.uint8)
            if totals[r]:
                dist.recv(buf, src=r)
            out[offs[r]: offs[r] + totals[r]] = buf.numpy()
        ret.put(bool(np.array_equal(out, O.encode(data))))
    elif mine.size:
        dist.send(torch.from_numpy(mine), dst=0)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 8192 * 9 + 77), (3, 8192 * 2), (2, 5000)])
def test_concatenated_shards_equal_whole_payload(world, n):
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get() is True


# ---- sharded decode: the segment hand-over protocol of gpuar_b200_decode_sharded, run on the CPU over gloo.
# Rank g owns segment g of the concatenated stream (S = gpuar_b200_shard_segment_bytes) plus a halo of the
# next one; it learns from rank g-1 where its first packet starts and how many packets precede it, walks the
# packets that START in its segment (the device does this with jump tables, index.cu), hands the exit on, and
# decodes its packets with the oracle.  The slices must tile the input exactly.
def _decode_worker(rank, world, port, n, ret):
    from gpuar_b200 import _lib
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data = D.mixed(13, n)
    pay = O.encode(data)                                      # the concatenated stream (what the ranks wrote)
    total = pay.size
    seg = int(_lib.lib().gpuar_b200_shard_segment_bytes(total, world))
    halo = 8704 + 512
    a, b = rank * seg, min(total, (rank + 1) * seg)
    mine = pay[a: min(total, b + halo)] if a < total else np.zeros(0, np.uint8)
    if rank == 0:
        entry, before, raw_before = 0, 0, 0
    else:
        msg = torch.zeros(3, dtype=torch.int64)
        dist.recv(msg, src=rank - 1)
        entry, before, raw_before = (int(v) for v in msg.tolist())
    pos, out, packets = entry, [], 0
    while pos < b:                                            # packets that start in [a, b)
        assert pos >= a
        loc = pos - a
        length = int(mine[loc]) | (int(mine[loc + 1]) << 8)
        assert loc + length <= mine.size                      # the halo covers the tail of the last packet
        out.append(O.decode(mine[loc: loc + length]))
        pos += length
        packets += 1
    got = np.concatenate(out) if out else np.zeros(0, np.uint8)
    if rank + 1 < world:
        dist.send(torch.tensor([pos, before + packets, raw_before + got.size], dtype=torch.int64), dst=rank + 1)
    ok = np.array_equal(got, data[raw_before: raw_before + got.size]) and raw_before == min(before * 8192, n)
    if rank + 1 == world:
        ok = ok and pos == total and raw_before + got.size == n
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret.put(bool(flag.item()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 8192 * 40 + 77), (3, 8192 * 9), (2, 5000), (3, 1 << 20)])
def test_segmented_decode_protocol(world, n):
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_decode_worker, args=(r, world, port, n, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get() is True
