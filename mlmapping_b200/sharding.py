"""Multi-GPU sharding of the naturally independent workloads: independent agent maps are
distributed round-robin over ranks and query streams are split evenly; there is no data-path
collective ("replicas only").  torch.distributed is plumbing: a barrier and the max-over-ranks /
sum-over-ranks reduction of the timing results.  Backend-agnostic so the logic is testable with gloo."""
from __future__ import annotations


def agents_for_rank(n_agents: int, rank: int, world: int) -> list[int]:
    """round-robin ownership of independent agent maps (CFG-D)"""
    return [a for a in range(n_agents) if a % world == rank]


def split_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """even contiguous split of a query stream: [begin, end) for this rank"""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def reduce_timing(times_ms: list[float], counts: list[float], device=None):
    """whole-job view of per-rank results: MAX over ranks of every time, SUM over ranks of every count"""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(times_ms), list(counts)
    t = torch.tensor(times_ms, dtype=torch.float64, device=device)
    c = torch.tensor(counts, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return t.tolist(), c.tolist()


def all_gather_blobs(blob: bytes, world: int, device=None) -> bytes:
    """the ranks' fixed-size setup blobs (mlm_shard_open / mlm_replica_open) concatenated in rank order: the one
    collective a sharded or replicated map needs, once, at setup.  `device`: where the backend wants its tensors
    (a CUDA device for nccl, None / cpu for gloo)"""
    import torch
    import torch.distributed as dist

    if world == 1:
        return bytes(blob)
    mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8)
    if device is not None:
        mine = mine.to(device)
    allb = torch.empty(world * len(blob), dtype=torch.uint8, device=mine.device)
    dist.all_gather_into_tensor(allb, mine)
    return bytes(allb.cpu().numpy().tobytes())
