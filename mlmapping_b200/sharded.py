"""One logical map sharded over the GPUs of a node (large LiDAR scans).

Every rank runs this class with the same scans.  Stage 1 is split by phi column, stage 2 by subbox owner;
in between the ranks all-gather the frame's distinct hit keys (so each derives the same libstdc++ iteration
order) and all-to-all the per-voxel update records.  torch.distributed (NCCL over NVLink) carries the two
exchanges; the compute on both sides is the CUDA library.  With world_size 1 (or no process group) the same
code path runs on one GPU, which is how the parity tests exercise it without a multi-GPU box."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .capi import FrameStats, MLMap, MlmConfig, _pose7


class _DevArray:
    """zero-copy view of library-owned device memory for torch (CUDA array interface, int32)"""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 2}


class ShardedMLMap:
    RECORD_INTS = 6  # 24-byte ShardRecord

    def __init__(self, cfg: MlmConfig, rank: int = 0, world: int = 1, device: int | None = None):
        import torch

        self.torch = torch
        self.rank, self.world = rank, world
        self.map = MLMap(cfg, device=rank if device is None else device)
        self.dev = torch.device("cuda", rank if device is None else device)
        self.last = {}
        self._pinned = None
        self._views = {}
        self.timing = True  # per-stage wall times in self.last (one device sync per stage); switch off for throughput runs

    def pinned_points(self, n: int) -> np.ndarray:
        """(n,3) float64 view of the page-locked scan buffer: fill it in place and pass it to integrate_points"""
        if self._pinned is None:
            self._pinned = self.map.pinned_array((max(n, self.map.cfg.max_points), 3), np.float64)
        return self._pinned[:n]

    def _join_torch_stream(self):
        """torch / NCCL work is enqueued on torch's current stream, the library runs on its own stream: results of
        the former must be complete before the library's kernels read them (the C ABI synchronises the other way)"""
        self.torch.cuda.current_stream(self.dev).synchronize()

    def _dist(self):
        import torch.distributed as dist
        return dist if (self.world > 1) else None

    def integrate_points(self, xyz: np.ndarray, T_wb) -> FrameStats:
        torch, m, lib = self.torch, self.map, self.map._lib
        dist = self._dist()
        import time
        tm = {}
        t_prev = time.perf_counter()

        def lap(name):
            nonlocal t_prev
            if not self.timing:
                return
            torch.cuda.synchronize(self.dev)
            now = time.perf_counter()
            tm[name] = round(1e6 * (now - t_prev), 1)
            t_prev = now
        pts = np.ascontiguousarray(np.asarray(xyz, dtype=np.float64).reshape(-1, 3))
        if self._pinned is None or self._pinned.shape[0] < pts.shape[0]:
            self._pinned = m.pinned_array((max(pts.shape[0], self.map.cfg.max_points), 3), np.float64)
        if pts.ctypes.data != self._pinned.ctypes.data:  # a scan produced in place (pinned_points) needs no host copy
            self._pinned[:pts.shape[0]] = pts  # page-locked staging: the H2D copy runs without a second host copy
        pts = self._pinned[:pts.shape[0]]
        n_hit, n_miss = C.c_int32(), C.c_int32()
        m._check(lib.mlm_shard_stage_points_f64(m._h, pts.ctypes.data, pts.shape[0], _pose7(T_wb), self.rank, self.world,
                                                C.byref(n_hit), C.byref(n_miss)))
        lap("stage_us")
        # ---- exchange 1: make the hit-map iteration order global ----
        n_total, fast = n_hit.value, False
        if dist:
            tot = torch.tensor([n_hit.value], dtype=torch.int64, device=self.dev)
            dist.all_reduce(tot)
            n_total = int(tot.item())
        act_ptr, B = C.c_void_p(), C.c_uint32()
        m._check(lib.mlm_shard_act_buffer(m._h, C.byref(act_ptr), C.byref(B)))
        if B.value > 1 and n_total <= B.value:
            # no rehash this frame: a key is cast by exactly one rank, so its stamp travels in its record and only
            # the bucket activation stamps need a min-all-reduce (unsigned order == int32 order after flipping bit 31)
            fast = True
            if dist:
                vk = (act_ptr.value, B.value)
                act = self._views.get(vk)
                if act is None:  # zero-copy torch view of the library's stamp array (one per parity / bucket count)
                    act = self._views[vk] = torch.as_tensor(_DevArray(act_ptr.value, B.value), device=self.dev)
                act.bitwise_xor_(-2 ** 31)
                dist.all_reduce(act, op=dist.ReduceOp.MIN)
                act.bitwise_xor_(-2 ** 31)
            lap("allgather_us")
            self._join_torch_stream()  # the min-all-reduce ran on torch's stream; the library's kernels must see its result
            m._check(lib.mlm_shard_order_fast(m._h, n_total))
        else:
            # rehash frame (map start / growth): gather every rank's (key, stamp) list; each rank re-sequences it
            keys = torch.empty(max(n_hit.value, 1), dtype=torch.int32, device=self.dev)
            stamps = torch.empty(max(n_hit.value, 1), dtype=torch.int32, device=self.dev)
            m._check(lib.mlm_shard_copy_hit_keys(m._h, keys.data_ptr(), stamps.data_ptr()))
            if dist:
                cnt = torch.tensor([n_hit.value], dtype=torch.int64, device=self.dev)
                all_cnt = torch.empty(self.world, dtype=torch.int64, device=self.dev)
                dist.all_gather_into_tensor(all_cnt, cnt)
                sizes = all_cnt.tolist()
                mx = max(max(sizes), 1)
                pad = torch.zeros((2, mx), dtype=torch.int32, device=self.dev)
                pad[0, :n_hit.value] = keys[:n_hit.value]
                pad[1, :n_hit.value] = stamps[:n_hit.value]
                g = torch.empty((self.world, 2, mx), dtype=torch.int32, device=self.dev)
                dist.all_gather_into_tensor(g, pad)
                keys_all = torch.cat([g[r, 0, :sizes[r]] for r in range(self.world)]).contiguous()
                stamps_all = torch.cat([g[r, 1, :sizes[r]] for r in range(self.world)]).contiguous()
            else:
                keys_all, stamps_all = keys[:n_hit.value].contiguous(), stamps[:n_hit.value].contiguous()
            assert int(keys_all.numel()) == n_total
            lap("allgather_us")
            self._join_torch_stream()  # gathered keys / stamps are consumed by kernels on the library's stream
            m._check(lib.mlm_shard_order(m._h, keys_all.data_ptr() if n_total else None,
                                         stamps_all.data_ptr() if n_total else None, n_total))
        lap("order_us")
        # ---- exchange 2: all-to-all of the per-voxel update records, grouped by owner ----
        send_counts = (C.c_int32 * self.world)()
        m._check(lib.mlm_shard_emit_counts(m._h, self.world, send_counts))
        sc = [int(v) for v in send_counts]
        send = torch.empty((max(sum(sc), 1), self.RECORD_INTS), dtype=torch.int32, device=self.dev)
        m._check(lib.mlm_shard_emit_pack(m._h, self.world, send_counts, send.data_ptr()))
        lap("emit_us")
        if dist:
            sct = torch.tensor(sc, dtype=torch.int64, device=self.dev)
            rct = torch.empty(self.world, dtype=torch.int64, device=self.dev)
            dist.all_to_all_single(rct, sct)
            rc = rct.tolist()
            recv = torch.empty((max(sum(rc), 1), self.RECORD_INTS), dtype=torch.int32, device=self.dev)
            dist.all_to_all_single(recv[:sum(rc)], send[:sum(sc)], output_split_sizes=rc, input_split_sizes=sc)
            n_recv = sum(rc)
        else:
            recv, n_recv = send, sum(sc)
        lap("alltoall_us")
        self._join_torch_stream()  # the received records are consumed by kernels on the library's stream
        st = FrameStats()
        m._check(lib.mlm_shard_ingest(m._h, recv.data_ptr() if n_recv else None, n_recv, C.byref(st)))
        lap("ingest_us")
        self.last = {"timing": tm, "fast_order": fast, "n_hit_local": n_hit.value, "n_hit_total": n_total, "n_miss_local": n_miss.value,
                     "records_sent": sum(sc), "records_received": n_recv,
                     "a2a_bytes": 24 * sum(sc), "order_exchange_bytes": 4 * B.value if fast else 8 * n_total}
        return st

    # queries / exports act on the subboxes this rank owns
    def export_map(self):
        return self.map.export_map()

    def close(self):
        self.map.close()


class ReplicatedMLMap:
    """One map updated on `src` rank and replicated on the others for split query streams: after each
    frame the dirty subbox blocks (those the frame touched) are broadcast and applied on the replicas."""

    def __init__(self, cfg: MlmConfig, rank: int = 0, world: int = 1, src: int = 0, device: int | None = None):
        import torch

        self.torch = torch
        self.rank, self.world, self.src = rank, world, src
        self.map = MLMap(cfg, device=rank if device is None else device)
        self.dev = torch.device("cuda", rank if device is None else device)
        self.last = {}

    def integrate_depth(self, img, T_wb):
        """call on every rank; only `src` needs the real image"""
        torch, m, lib = self.torch, self.map, self.map._lib
        st = None
        n, rb = C.c_int32(0), C.c_size_t(0)
        if self.rank == self.src:
            st = m.integrate_depth(img, T_wb)
            m._check(lib.mlm_dirty_count(m._h, C.byref(n), C.byref(rb)))
        if self.world > 1:
            import torch.distributed as dist
            meta = torch.tensor([n.value, rb.value], dtype=torch.int64, device=self.dev)
            dist.broadcast(meta, src=self.src)
            nb, rbytes = int(meta[0]), int(meta[1])
        else:
            nb, rbytes = n.value, rb.value
        buf = torch.empty(max(nb * rbytes, 16), dtype=torch.uint8, device=self.dev)
        if self.rank == self.src:
            m._check(lib.mlm_dirty_export(m._h, buf.data_ptr(), nb))
        if self.world > 1:
            import torch.distributed as dist
            dist.broadcast(buf, src=self.src)
            torch.cuda.synchronize(self.dev)
        if self.rank != self.src:
            m._check(lib.mlm_dirty_import(m._h, buf.data_ptr(), nb))
        self.last = {"dirty_blocks": nb, "broadcast_bytes": nb * rbytes}
        return st

    def import_from(self, buf_ptr: int, nb: int):
        self.map._check(self.map._lib.mlm_dirty_import(self.map._h, buf_ptr, nb))
