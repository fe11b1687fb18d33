"""One logical map over the GPUs of a node: thin callers of the C ABI.

ShardedMLMap — large LiDAR scans: every rank gets the same scan, casts its phi columns and owns the subboxes whose
index hashes to it.  The two exchanges (all-gather of the hit keys, all-to-all of the update records) happen INSIDE
the library, as stores into the owners' exchange arenas over NVLink peer memory (mlm_shard_*, csrc/shard_kernels.cuh);
this module only carries the 128-byte setup blobs between the ranks once (torch.distributed all_gather or, for
several ranks inside one process, a plain list) and then calls submit / finish per scan.

ReplicatedMLMap — one map updated on `src` and replicated for split query streams: the dirty subbox blocks of each
frame are broadcast (NCCL) and applied on the replicas."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .capi import SHARD_BLOB_BYTES, FrameStats, MLMap, MlmConfig, ShardExchange, _pose7
from .sharding import all_gather_blobs


class ShardedMLMap:
    def __init__(self, cfg: MlmConfig, rank: int = 0, world: int = 1, device: int | None = None, connect: bool = True):
        self.rank, self.world = rank, world
        self.map = MLMap(cfg, device=rank if device is None else device)
        self._pinned = None
        self.last = {}
        blob = (C.c_ubyte * SHARD_BLOB_BYTES)()
        self.map._check(self.map._lib.mlm_shard_open(self.map._h, rank, world, blob))
        self.blob = bytes(blob)
        if connect and world > 1:
            self.connect(self._all_gather_blobs())

    def _all_gather_blobs(self):
        import torch
        import torch.distributed as dist
        dev = torch.device("cuda", self.map.device) if dist.get_backend() == "nccl" else None
        return all_gather_blobs(self.blob, self.world, dev)

    def connect(self, blobs: bytes):
        """blobs: the `world` setup blobs in rank order"""
        assert len(blobs) == self.world * SHARD_BLOB_BYTES
        buf = (C.c_ubyte * len(blobs)).from_buffer_copy(blobs)
        self.map._check(self.map._lib.mlm_shard_connect(self.map._h, buf))

    def pinned_points(self, n: int) -> np.ndarray:
        """(n,3) float64 view of a page-locked scan buffer: fill it in place and pass it to submit / integrate_points"""
        if self._pinned is None or self._pinned.shape[0] < n:
            self._pinned = self.map.pinned_array((max(n, self.map.cfg.max_points), 3), np.float64)
        return self._pinned[:n]

    def submit(self, xyz: np.ndarray, T_wb):
        """enqueue one scan (H2D copy, staging, exchange, owner-side fusion) on the library's stream; no host wait"""
        pts = np.ascontiguousarray(np.asarray(xyz, dtype=np.float64).reshape(-1, 3))
        m = self.map
        m._check(m._lib.mlm_shard_submit_points_f64(m._h, pts.ctypes.data, pts.shape[0], _pose7(T_wb)))
        self._keep = pts

    def submit_slice(self, xyz_slice: np.ndarray, first: int, n_total: int, T_wb):
        """this rank's slice [first, first + len(xyz_slice)) of a scan of n_total points (the ranks' slices partition it):
        1/world of the host-to-device traffic per rank, the rest travels over NVLink inside the library"""
        pts = np.ascontiguousarray(np.asarray(xyz_slice, dtype=np.float64).reshape(-1, 3))
        m = self.map
        m._check(m._lib.mlm_shard_submit_points_slice_f64(m._h, pts.ctypes.data, first, pts.shape[0], n_total, _pose7(T_wb)))
        self._keep = pts

    def submit_device(self, d_xyz: int, n: int, T_wb):
        m = self.map
        m._check(m._lib.mlm_shard_submit_points_f64_device(m._h, d_xyz, n, _pose7(T_wb)))

    def finish(self) -> FrameStats:
        m = self.map
        st = FrameStats()
        m._check(m._lib.mlm_shard_finish(m._h, C.byref(st)))
        ex = ShardExchange()
        m._check(m._lib.mlm_shard_last_exchange(m._h, C.byref(ex)))
        self.last = ex.as_dict()
        self.last["a2a_bytes_in"] = 24 * ex.records_received
        self.last["gather_bytes_in"] = 8 * ex.n_hit_total
        return st

    STAGES = ("resets_h2d", "k_project", "k_column", "k_shard_push", "k_shard_act_ingest", "k_fuse")

    def last_kernel_us(self):
        """per-stage device times of the last scan (needs self.map.set_profiling(True))"""
        ms = (C.c_float * len(self.STAGES))()
        self.map._check(self.map._lib.mlm_shard_last_kernel_ms(self.map._h, ms))
        return {k: 1e3 * float(v) for k, v in zip(self.STAGES, ms)}

    def integrate_points(self, xyz: np.ndarray, T_wb) -> FrameStats:
        self.submit(xyz, T_wb)
        return self.finish()

    # queries / exports act on the subboxes this rank owns
    def export_map(self):
        return self.map.export_map()

    def close(self):
        if self.map._h is not None and self.map._h.value:
            self.map._lib.mlm_shard_close(self.map._h)
        self.map.close()


def sharded_group_in_process(cfg: MlmConfig, world: int, devices=None):
    """`world` ranks driven by ONE process (several GPUs, or several ranks on one GPU as the single-GPU tests do):
    the blobs travel through a list.  Submit on all ranks before finishing any."""
    devices = devices or [0] * world
    ranks = [ShardedMLMap(cfg, rank=r, world=world, device=devices[r], connect=False) for r in range(world)]
    blobs = b"".join(r.blob for r in ranks)
    if world > 1:
        for r in ranks:
            r.connect(blobs)
    return ranks


class ReplicatedMLMap:
    """One map updated on `src` rank and replicated on the others for split query streams: after each
    frame the dirty subbox blocks (those the frame touched) reach the replicas.

    transport "p2p" (default): the library's kernels store the records straight into the replicas' inboxes over NVLink
    peer memory (mlm_replica_*); this class only carries the setup blobs once.  transport "nccl": the caller-side
    broadcast of mlm_dirty_export / mlm_dirty_import buffers (kept as the baseline and for devices without peer access)."""

    def __init__(self, cfg: MlmConfig, rank: int = 0, world: int = 1, src: int = 0, device: int | None = None,
                 transport: str = "p2p", connect: bool = True):
        import torch

        self.torch = torch
        self.rank, self.world, self.src = rank, world, src
        self.transport = transport
        self.map = MLMap(cfg, device=rank if device is None else device)
        self.dev = torch.device("cuda", rank if device is None else device)
        self.last = {}
        self.blob = b""
        if transport == "p2p":
            blob = (C.c_ubyte * SHARD_BLOB_BYTES)()
            self.map._check(self.map._lib.mlm_replica_open(self.map._h, rank, world, src, blob))
            self.blob = bytes(blob)
            if connect and world > 1:
                self.connect(self._all_gather_blobs())

    def _all_gather_blobs(self):
        import torch.distributed as dist
        return all_gather_blobs(self.blob, self.world, self.dev if dist.get_backend() == "nccl" else None)

    def connect(self, blobs: bytes):
        assert len(blobs) == self.world * SHARD_BLOB_BYTES
        buf = (C.c_ubyte * len(blobs)).from_buffer_copy(blobs)
        self.map._check(self.map._lib.mlm_replica_connect(self.map._h, buf))

    def integrate_depth(self, img, T_wb):
        """call on every rank; only `src` needs the real image"""
        if self.transport != "p2p":
            return self._integrate_depth_nccl(img, T_wb)
        m, lib = self.map, self.map._lib
        st = None
        n = C.c_int32(0)
        if self.rank == self.src:
            st = m.integrate_depth(img, T_wb)
            if self.world > 1:
                m._check(lib.mlm_replica_publish(m._h, C.byref(n)))
            else:
                n.value = 0
        else:
            m._check(lib.mlm_replica_apply(m._h, C.byref(n)))
        self.last = {"dirty_blocks": n.value, "broadcast_bytes": n.value * ((16 + 6 * m.cells + 15) & ~15)}
        return st

    def close(self):
        if self.map._h is not None and self.map._h.value:
            if self.transport == "p2p":
                self.map._lib.mlm_replica_close(self.map._h)
        self.map.close()

    def _integrate_depth_nccl(self, img, T_wb):
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


def replicated_group_in_process(cfg: MlmConfig, world: int, src: int = 0, devices=None):
    """`world` ranks of a replicated map driven by ONE process (several GPUs, or several handles on one GPU as the
    single-GPU tests do).  Per frame: integrate_depth on the source first, then on every replica."""
    devices = devices or [0] * world
    ranks = [ReplicatedMLMap(cfg, rank=r, world=world, src=src, device=devices[r], connect=False) for r in range(world)]
    blobs = b"".join(r.blob for r in ranks)
    if world > 1:
        for r in ranks:
            r.connect(blobs)
    return ranks
