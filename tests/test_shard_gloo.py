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
