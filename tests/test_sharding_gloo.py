"""CPU, world_size 2 over gloo: the N>1 plumbing of bench.py (agent ownership, query split, max/sum reduce)"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mlmapping_b200.sharding import agents_for_rank, all_gather_blobs, reduce_timing, split_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    agents = agents_for_rank(8, rank, world)
    b, e = split_range(10_000_001, rank, world)
    # each rank "measures" its own time and ray count
    times, counts = reduce_timing([10.0 + rank, 5.0 - rank], [1000.0 * (rank + 1), float(len(agents))])
    # the setup blobs of a sharded / replicated map travel once, in rank order
    blobs = all_gather_blobs(bytes([rank + 1]) * 128, world)
    dist.barrier()
    out[rank] = (agents, (b, e), times, counts, blobs)
    dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    a0, r0, t0, c0, b0 = out[0]
    a1, r1, t1, c1, b1 = out[1]
    assert b0 == b1 == bytes([1]) * 128 + bytes([2]) * 128
    assert sorted(a0 + a1) == list(range(8)) and not set(a0) & set(a1)
    assert r0[0] == 0 and r0[1] == r1[0] and r1[1] == 10_000_001
    assert t0 == t1 == [11.0, 5.0]          # max over ranks
    assert c0 == c1 == [3000.0, 8.0]        # sum over ranks


def test_single_process_passthrough():
    assert reduce_timing([1.0], [2.0]) == ([1.0], [2.0])
    assert agents_for_rank(8, 0, 1) == list(range(8))
    assert split_range(10, 2, 3) == (7, 10)
    assert all_gather_blobs(b"x" * 128, 1) == b"x" * 128
    # scan slices of a sharded map (bench.py): the ranks' ranges partition the scan, also when it does not divide evenly
    for n, world in ((262144, 8), (7, 3), (2, 4), (0, 2)):
        cuts = [split_range(n, r, world) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n and all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))
